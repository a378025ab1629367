// attn_bwd.cuh — K2 / K3 backward: persistent flash-attention backward on tcgen05 (included by attn.cu).
//
//   prep kernel: delta[b,h,i] = rowsum(dO * O)
//   main kernel: persistent, one CTA per SM; a work item = (head, 128-key tile, batch), items of a CTA are contiguous
//   in that order, so the relative-bias window and its gradient bins stay in shared memory across ~15 items and the
//   K / V loads, TMEM allocation and barrier set-up of an item hide under the previous item's tail.  Per item a loop
//   over 128-query tiles with a 2-stage TMA ring of Q / dO:
//     S = Q K^T, dP = dO V^T                  (TMEM [0,128) and [128,256))
//     P = exp2(s - lse), dS = P * (dP - delta) * scale   -> bf16 [query][key] tiles in smem
//     dV += P^T dO, dK += dS^T Q               (A operands read MN-major from those tiles; TMEM [256,320), [320,384))
//     dQ_m = dS K -> TMEM [384,448) / [448,512) by tile parity -> fp32 staging tile -> ONE TMA reduce-add
//   16 compute warps (TMEM lane = query row, a quarter of the 128 key columns per thread) + 1 issuer warp that owns
//   TMA and tcgen05.mma: S/dP of tile g+1 are issued as soon as every compute thread holds tile g in registers, the
//   three accumulating GEMMs of tile g as soon as its P/dS are in smem.
// Instruction budget of the compute warps (issue slots per score, relative bias + dropout on; the round-1 kernel
// spent ~46): 0.75 LDS, 1.5 packed bias / score / -lse adds, 1 MUFU.EX2, 1.5 packed dS, 1 F2FP, 2 dropout masking
// (PRMT sign-replicate + AND on the fp32 P), ~2.3 Philox + bit-sliced compare, 3 d_rel (one shuffle per column, two
// predicated adds), 0.25 STS, ~2 per-tile overhead = ~15.
#pragma once

namespace pvqa {

constexpr int kBComputeWarps = 16;
constexpr int kBComputeThreads = kBComputeWarps * 32;
constexpr int kBSyncThreads = kBComputeThreads + 32;      // named barriers 1 and 3: compute warps + the issuer warp
constexpr int kBThreads = kBSyncThreads;                  // (ptxas fits the compute path into the 544-thread budget
                                                          //  without spills; setmaxnreg splits only made it spill)
constexpr uint32_t kBTmemCols = 512;   // S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ [384,448), [448,512)
constexpr int kBTile = kBN * kD * 2;                      // 16 KB: a 128-row bf16 tile
constexpr int kBOffK = 0;                                 // 2 x 16 KB: K of item n in buffer n & 1
constexpr int kBOffV = kBOffK + 2 * kBTile;               // 16 KB (free again once an item's last dP is done)
constexpr int kBOffQ = kBOffV + kBTile;                   // 2 stages x (Q 16 KB, dO 16 KB)
constexpr int kBOffP = kBOffQ + 4 * kBTile;               // 32 KB
constexpr int kBOffdS = kBOffP + kBM * kBN * 2;           // 32 KB
constexpr int kBOffStg = kBOffdS + kBM * kBN * 2;         // 32 KB: fp32 dQ staging, two [128][32] SW128 halves
constexpr int kBOffBar = kBOffStg + kBM * kD * 4;         // 208 KB
constexpr int kBOffFloats = kBOffBar + 128;               // kadd[2][128], relc[2][cs], drel[n_win], scp[32], dscp[16][32]

__host__ __device__ constexpr int b_rel_copy_stride(int n_qpad) { return ((n_qpad + kBN + 2 + 31) / 32) * 32 + 16; }

struct AttnBwdParams {
  const float* lse;
  const float* delta;         // (B,H,Sq)
  const float* rel_bias;
  const float* key_add;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  float* d_rel;               // (H, Sq+Sk-1) fp32 accumulated, or null
  int B, H, Sq, Sk;
  long long dk_stride_b, dk_stride_s, dk_stride_h;
  long long dv_stride_b, dv_stride_s, dv_stride_h;
  float scale, sl2;           // sl2 = scale * log2(e)
  AttnDrop drop;
  const uint8_t* scp_bucket;  // SaL SCP bias (see AttnFwdParams)
  const float* scp_tab;
  float* d_scp;               // (H, 32) fp32 accumulated, or null
  int scp_q0, scp_L;
  int n_qt, n_kt, n_items;    // query tiles, key tiles, H * n_kt * B
  int rel_far;                // |j - i| >= rel_far => the bias depends only on sign(j - i)  (0: no such promise)
};

struct AttnPrepParams {
  const __nv_bfloat16* o; const __nv_bfloat16* d_o; float* delta;
  int B, H, Sq;
  long long o_stride_b, o_stride_s, o_stride_h, do_stride_b, do_stride_s, do_stride_h;
};

__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const AttnPrepParams p) {
  // 8 lanes per (b, i, h) row of 64 elements, rows enumerated in memory order (h fastest): every warp reads
  // four whole 128-byte rows of O and of dO per step
  const long long n = (long long)p.B * p.H * p.Sq;
  const int sub = threadIdx.x & 7;
  // warp-uniform trip count (the shuffles below need all 32 lanes): t0 = first row of this warp's group of four
  for (long long t0 = ((long long)blockIdx.x * 256 + (threadIdx.x & ~31)) >> 3; t0 < n; t0 += ((long long)gridDim.x * 256) >> 3) {
    const long long t = t0 + ((threadIdx.x & 31) >> 3);
    const bool valid = t < n;
    const long long tc = valid ? t : 0;
    const int h = (int)(tc % p.H);
    const int i = (int)((tc / p.H) % p.Sq);
    const int b = (int)(tc / ((long long)p.H * p.Sq));
    const __nv_bfloat16* orow = p.o + b * p.o_stride_b + i * p.o_stride_s + h * p.o_stride_h + sub * 8;
    const __nv_bfloat16* grow = p.d_o + b * p.do_stride_b + i * p.do_stride_s + h * p.do_stride_h + sub * 8;
    const f8 a = Vec8<__nv_bfloat16>::load(orow);
    const f8 g = Vec8<__nv_bfloat16>::load(grow);
    float acc = 0.f;
#pragma unroll
    for (int x = 0; x < 8; ++x) acc = fmaf(a.v[x], g.v[x], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && valid) p.delta[((long long)b * p.H + h) * p.Sq + i] = acc;
  }
}

#ifdef PVQA_ATTN_TRACE
#define PVQA_TRACEB(ev)                                                                             \
  do {                                                                                              \
    if (blockIdx.x < 64 && (ev) < 32) {                                                             \
      if (threadIdx.x == 0) g_attn_trace[blockIdx.x * 64 + (ev)] = clock64();                       \
      if (threadIdx.x == kBComputeThreads) g_attn_trace[blockIdx.x * 64 + 32 + (ev)] = clock64();   \
    }                                                                                               \
  } while (0)
#else
#define PVQA_TRACEB(ev)
#endif

template <bool HAS_REL, bool DROP, bool CAUSAL, bool SCP>
__global__ void __launch_bounds__(kBThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an OFFSET into the __shared__ array so that the compiler
  // keeps the shared address space (32-bit LDS/STS instead of generic 64-bit LD/ST for every smem access)
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_k = reinterpret_cast<uint64_t*>(smem + kBOffBar);   // [2] K of item n landed, buffer n & 1
  uint64_t* bar_v = bar_k + 2;            // V of an item landed
  uint64_t* bar_ld = bar_k + 3;           // [2] Q / dO of tile g landed, stage g & 1
  uint64_t* bar_s = bar_k + 5;            // S/dP of a query tile are in TMEM                       (tcgen05.commit)
  uint64_t* bar_g = bar_k + 6;            // the three accumulating GEMMs of a query tile are done  (tcgen05.commit)
  uint64_t* bar_stg = bar_k + 7;          // the dQ staging tile has been read by its reduce
  uint64_t* bar_r = bar_k + 8;            // dV / dK GEMMs of a query tile are done: its Q / dO ring stage is free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_k + 9);
  const int n_rel = p.Sq + p.Sk - 1;
  const int n_qpad = p.n_qt * kBM;
  const int n_win = n_qpad + kBN;                                // window of relative offsets one key tile can see
  const int cs = b_rel_copy_stride(n_qpad);
  float* s_kadd = reinterpret_cast<float*>(smem + kBOffFloats);  // [2][kBN] key term of item n in buffer n & 1
  float* s_relc = s_kadd + 2 * kBN;                              // [2][cs] two shifted copies of the bias window
  float* s_drel = s_relc + (HAS_REL ? 2 * cs : 0);               // [n_win] gradient bins of the window
  float* s_scp = s_drel + (HAS_REL ? n_win : 0);                 // [32] SCP table of this head
  float* s_dscp = s_scp + 32;                                    // [16 warps][32] SCP gradient bins

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer_wg = warp >= kBComputeWarps;
  const bool is_issuer = warp == kBComputeWarps;        // warp 16: TMA, tcgen05.mma and the dQ reduce, nothing else
  // this CTA's contiguous range of items; item w = (h * n_kt + kt) * B + b
  const int w0 = (int)((long long)blockIdx.x * p.n_items / gridDim.x);
  const int w1 = (int)((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
  auto decode = [&](int w, int& h, int& kt, int& b) {
    b = w % p.B;
    const int hk = w / p.B;
    kt = hk % p.n_kt;
    h = hk / p.n_kt;
  };
  // items are walked incrementally (b fastest, then key tile, then head): integer division by a run-time divisor costs
  // ~150 cycles, and the first flat version paid a dozen of them per tile
  struct Item { int h, kt, b; };
  auto item_of = [&](int w) { Item x; decode(w, x.h, x.kt, x.b); return x; };
  auto next_item = [&](Item x) {
    if (++x.b == p.B) { x.b = 0; if (++x.kt == p.n_kt) { x.kt = 0; ++x.h; } }
    return x;
  };
  // query tiles an item visits: [m_first, n_qt); tiles entirely above the diagonal see nothing
  auto first_tile = [&](int kt) { return CAUSAL ? kt : 0; };        // (kBN == kBM)
  PVQA_TRACEB(0);

  if (is_issuer) {
    if (lane == 0) {
      tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV); tc05::prefetch_tmap(&tmdO);
      tc05::prefetch_tmap(&tmdQ);
      for (int x = 0; x < 9; ++x) tc05::mbar_init(bar_k + x, 1);
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_slot, kBTmemCols);
    tc05::tmem_relinquish();
  }
  // ---- staging (compute threads); everything additive is pre-multiplied by log2(e) ----
  // staged[x] = bias of relative index r = (j0 + Sq - n_qpad) + x, x in [0, n_win): the score of query i and local
  // key jl reads x = jl + (n_qpad - 1 - i).  Two copies: copy_k[a] = staged[a + k]  (8-byte loads, see attn_fwd.cuh)
  auto stage_rel = [&](int h, int kt) {
    const int base_r = kt * kBN + p.Sq - n_qpad;
    for (int x = tid; x < 2 * cs; x += kBComputeThreads) {
      const int k = x >= cs ? 1 : 0;
      const int y = x - k * cs + k;
      const int r = base_r + y;
      s_relc[x] = (y < n_win && r >= 0 && r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
    }
    for (int x = tid; x < n_win; x += kBComputeThreads) s_drel[x] = 0.f;
    if (SCP) {
      if (tid < 32) s_scp[tid] = p.scp_tab[h * 32 + tid] * kLog2e;
      s_dscp[tid] = 0.f;                       // 512 threads == 16 x 32 bins
    }
  };
  // bins -> global gradient (the smem tile holds scale * dS)
  auto flush_rel = [&](int h, int kt) {
    const float inv_scale = 1.0f / p.scale;
    if (p.d_rel) {
      const int base_r = kt * kBN + p.Sq - n_qpad;
      for (int x = tid; x < n_win; x += kBComputeThreads) {
        const int r = base_r + x;
        const float gsum = s_drel[x];
        if (r >= 0 && r < n_rel && gsum != 0.f) atomicAdd(p.d_rel + (long long)h * n_rel + r, gsum * inv_scale);
      }
    }
    if (SCP && p.d_scp && tid < 32) {
      float gsum = 0.f;
#pragma unroll
      for (int w = 0; w < kBComputeWarps; ++w) gsum += s_dscp[w * 32 + tid];
      if (gsum != 0.f) atomicAdd(p.d_scp + h * 32 + tid, gsum * inv_scale);
    }
  };
  if (!is_issuer_wg && w0 < w1) {
    int h, kt, b;
    decode(w0, h, kt, b);
    if (tid < kBN) {
      const int j = kt * kBN + tid;
      s_kadd[tid] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
    }
    if (HAS_REL) stage_rel(h, kt);
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  PVQA_TRACEB(1);

  if (is_issuer_wg) {
    if (is_issuer && w0 < w1) {
      // ================= issuer warp.  Every lane walks the tile sequence and waits on the barriers; one elected lane
      // issues TMA / tcgen05.  Everything is warp-uniform, so addresses and descriptors live in uniform registers
      // (under `if (lane == 0)` ptxas moved each tcgen05.mma operand through an R2UR waterfall: ~100 cycles per MMA,
      // 3 000 per tile for the 32 MMAs — profiles/r02_call3_attn_phase_trace.txt) =================
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc_s = tc05::idesc_bf16(kBM, kBN, 0, 0);
      const uint32_t idesc_dkv = tc05::idesc_bf16(kBN, kD, 1, 1);   // A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      const uint32_t idesc_dq = tc05::idesc_bf16(kBM, kD, 0, 1);    // A = dS (K-major), B = K (MN-major)
      const uint32_t smem0 = tc05::smem_u32(smem);
      struct Cursor { int w, it; Item x; };                // a (item, query tile) position in this CTA's sequence
      auto enter = [&](Cursor& c, int w, Item x) { c.w = w; c.x = x; c.it = first_tile(x.kt); };
      auto advance = [&](Cursor& c) {
        if (++c.it < p.n_qt) return true;
        if (c.w + 1 >= w1) return false;
        enter(c, c.w + 1, next_item(c.x));
        return true;
      };
      auto is_first = [&](const Cursor& c) { return c.it == first_tile(c.x.kt); };
      auto load_k = [&](const Cursor& c) {
        const int h = c.x.h, kt = c.x.kt, b = c.x.b;
        const int nb = (c.w - w0) & 1;
        if (tc05::elect_one()) {
          tc05::mbar_expect_tx(bar_k + nb, kBTile);
          tc05::tma_load_4d(smem + kBOffK + nb * kBTile, &tmK, bar_k + nb, 0, h, kt * kBN, b);
        }
      };
      auto load_v = [&](const Cursor& c) {
        const int h = c.x.h, kt = c.x.kt, b = c.x.b;
        if (tc05::elect_one()) {
          tc05::mbar_expect_tx(bar_v, kBTile);
          tc05::tma_load_4d(smem + kBOffV, &tmV, bar_v, 0, h, kt * kBN, b);
        }
      };
      auto load_qdo = [&](const Cursor& c, int gx) {
        const int h = c.x.h, b = c.x.b;
        uint8_t* dst = smem + kBOffQ + (gx & 1) * (2 * kBTile);
        if (tc05::elect_one()) {
          tc05::mbar_expect_tx(bar_ld + (gx & 1), 2 * kBTile);
          tc05::tma_load_4d(dst, &tmQ, bar_ld + (gx & 1), 0, h, c.it * kBM, b);
          tc05::tma_load_4d(dst + kBTile, &tmdO, bar_ld + (gx & 1), 0, h, c.it * kBM, b);
        }
      };
      auto issue_s = [&](const Cursor& c, int gx) {         // S = Q K^T of tile gx
        const uint64_t qd_ = tc05::desc_sw128_k(smem0 + kBOffQ + (gx & 1) * (2 * kBTile));
        const uint64_t kd_ = tc05::desc_sw128_k(smem0 + kBOffK + ((c.w - w0) & 1) * kBTile);
        if (tc05::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)
            tc05::mma_bf16_ss(tmem_u, tc05::desc_step(qd_, ks * 32), tc05::desc_step(kd_, ks * 32), idesc_s, ks > 0);
        }
      };
      auto issue_dp = [&](int gx) {                         // dP = dO V^T of tile gx, then "S/dP ready"
        const uint64_t dd_ = tc05::desc_sw128_k(smem0 + kBOffQ + (gx & 1) * (2 * kBTile) + kBTile);
        const uint64_t vd_ = tc05::desc_sw128_k(smem0 + kBOffV);
        if (tc05::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)
            tc05::mma_bf16_ss(tmem_u + kBN, tc05::desc_step(dd_, ks * 32), tc05::desc_step(vd_, ks * 32), idesc_s, ks > 0);
          tc05::mma_commit(bar_s);
        }
      };
      auto reduce_dq = [&](const Cursor& c) {   // staging tile (fp32, two [128][32] SW128 halves) += into dq_accum
        const int h = c.x.h, b = c.x.b;
        if (tc05::elect_one()) {
          tc05::tma_reduce_add_4d(&tmdQ, smem + kBOffStg, 0, h, c.it * kBM, b);
          tc05::tma_reduce_add_4d(&tmdQ, smem + kBOffStg + kBM * 128, 32, h, c.it * kBM, b);
          tc05::bulk_commit_group();
        }
      };
      auto release_stg = [&]() {                // the reduce has read the staging tile: hand it back (same elected lane:
        if (tc05::elect_one()) {                //  bulk groups belong to the thread that committed them)
          tc05::bulk_wait_group_read0();
          tc05::mbar_arrive(bar_stg);
        }
      };

      Cursor c_g, c_s, c_ld, c_prev;            // tiles g, g + 1, g + 2 and g - 1
      enter(c_g, w0, item_of(w0));
      c_s = c_g; c_ld = c_g; c_prev = c_g;
      int items_k = 1;                          // items whose K load has been issued
      int n_v = 1;                              // V loads issued (one per item: the phase of bar_v)
      load_k(c_g); load_v(c_g); load_qdo(c_ld, 0);
      bool has_ld = advance(c_ld);
      if (has_ld) {
        if (c_ld.w != c_g.w) {                  // (the V buffer is still the first item's)
          load_k(c_ld);
          items_k = 2;
        }
        load_qdo(c_ld, 1);
      }
      bool has_s = has_ld;
      if (has_ld) has_ld = advance(c_ld);
      tc05::mbar_wait(bar_k, 0);
      tc05::mbar_wait(bar_v, 0);
      tc05::mbar_wait(bar_ld, 0);
      tc05::tc_fence_after_sync();
      issue_s(c_s, 0);
      issue_dp(0);
      if (has_s) advance(c_s);
      for (int g = 0;; ++g) {
        const bool last_of_item = c_g.it + 1 == p.n_qt;
        const bool first = is_first(c_g);
        const bool tr = g < 3;
        if (tr) PVQA_TRACEB(2 + 6 * g);
        // (1) every compute thread holds S/dP of tile g in registers: TMEM S/dP can take tile g + 1.  When tile g + 1 opens
        //     the next item, its V is fetched into the buffer that just became free and dP waits for it at the end of
        //     this turn; S needs only K, which arrived an item ago.
        tc05::named_bar_sync(1, kBSyncThreads);
        if (tr) PVQA_TRACEB(3 + 6 * g);
        bool dp_pending = false;
        if (has_s) {
          tc05::tc_fence_after_sync();
          if (last_of_item) {
            load_v(c_s);
            tc05::mbar_wait(bar_k + ((c_s.w - w0) & 1), ((c_s.w - w0) >> 1) & 1);
            ++n_v;
            dp_pending = true;
          }
          tc05::mbar_wait(bar_ld + ((g + 1) & 1), ((g + 1) >> 1) & 1);
          tc05::tc_fence_after_sync();
          issue_s(c_s, g + 1);
          if (!last_of_item) issue_dp(g + 1);
          has_s = advance(c_s);
        }
        if (tr) PVQA_TRACEB(4 + 6 * g);
        // (2) P/dS of tile g are in smem, dQ of tile g-1 is staged and, when tile g opens an item, the compute warps have
        //     drained the previous item's dK / dV
        tc05::named_bar_sync(3, kBSyncThreads);
        if (tr) PVQA_TRACEB(5 + 6 * g);
        tc05::tc_fence_after_sync();
        {
          const uint32_t q_addr = smem0 + kBOffQ + (g & 1) * (2 * kBTile);
          const uint64_t qd_ = tc05::desc_sw128_k(q_addr), dod_ = tc05::desc_sw128_k(q_addr + kBTile);
          const uint64_t kd_ = tc05::desc_sw128_k(smem0 + kBOffK + ((c_g.w - w0) & 1) * kBTile);
          const uint64_t pt_ = tc05::smem_desc_sw128(smem0 + kBOffP, kBM * 128, 1024);      // P^T: MN-major A
          const uint64_t dst_ = tc05::smem_desc_sw128(smem0 + kBOffdS, kBM * 128, 1024);    // dS^T
          const uint64_t dsd_ = tc05::desc_sw128_k(smem0 + kBOffdS);                        // dS: K-major A
          const uint32_t dq_col = tmem_u + 384 + (g & 1) * 64;
          const uint32_t acc0 = first ? 0u : 1u;
          if (tc05::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kBM / 16; ++ks)     // dV += P^T dO   (K = 128 query rows, 16 per step)
              tc05::mma_bf16_ss(tmem_u + 256, tc05::desc_step(pt_, ks * 2048), tc05::desc_step(dod_, ks * 2048), idesc_dkv,
                                ks > 0 ? 1u : acc0);
#pragma unroll
            for (int ks = 0; ks < kBM / 16; ++ks)     // dK += dS^T Q
              tc05::mma_bf16_ss(tmem_u + 320, tc05::desc_step(dst_, ks * 2048), tc05::desc_step(qd_, ks * 2048), idesc_dkv,
                                ks > 0 ? 1u : acc0);
            tc05::mma_commit(bar_r);                  // Q / dO of this tile have been consumed: the ring stage is free
#pragma unroll
            for (int ks = 0; ks < kBN / 16; ++ks)     // dQ_m = dS K    (K = 128 keys)
              tc05::mma_bf16_ss(dq_col, tc05::desc_step(dsd_, (ks >> 2) * (kBM * 128) + (ks & 3) * 32),
                                tc05::desc_step(kd_, ks * 2048), idesc_dq, ks > 0);
            tc05::mma_commit(bar_g);
          }
          if (g > 0) { reduce_dq(c_prev); release_stg(); }
        }
        if (tr) PVQA_TRACEB(6 + 6 * g);
        // the ring stage of tile g takes tile g + 2 as soon as dV / dK of tile g are done; a new item's K goes out with
        // its first Q / dO (into the buffer of the item before the current one, whose GEMMs finished long ago)
        if (has_ld) {
          const bool need_k = c_ld.w - w0 >= items_k;
          tc05::mbar_wait(bar_r, g & 1);
          load_qdo(c_ld, g + 2);
          if (need_k) {                             // the K buffer may have fed dQ of tile g (single-tile items)
            tc05::mbar_wait(bar_g, g & 1);
            load_k(c_ld);
            ++items_k;
          }
          has_ld = advance(c_ld);
        }
        if (dp_pending) {                           // the next item's V has had a whole tile of time to land
          tc05::mbar_wait(bar_v, (n_v - 1) & 1);
          tc05::tc_fence_after_sync();
          issue_dp(g + 1);
        }
        c_prev = c_g;
        if (tr) PVQA_TRACEB(7 + 6 * g);
        if (!advance(c_g)) break;
      }
      // dQ of the sequence's last tile is staged
      tc05::named_bar_sync(3, kBSyncThreads);
      reduce_dq(c_prev);
      if (tc05::elect_one()) tc05::bulk_wait_group0();      // all reductions performed before the CTA retires
      __syncwarp();
    }
  }
  // (compute warps continue below; the issuer warpgroup joins them at the final barrier)
  if (!is_issuer_wg && w0 < w1) {
    // ================= 16 compute warps =================
    const int quad = warp & 3;                   // TMEM lane quadrant == 32-row group of the query tile
    const int qd = warp >> 2;                    // which quarter of the key columns this thread owns
    const int rowl = quad * 32 + lane;           // row inside the 128-row tile == TMEM lane
    const int jl0 = qd * 32;                     // first local key column of this thread
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint64_t rng_off = p.drop.offset + ((DROP && p.drop.rng_base) ? *p.drop.rng_base : 0ull);
    // this thread's 16-byte chunks of the P / dS rows with the 128-byte swizzle folded in (chunk c: ^ (c << 4))
    const uint32_t row_x = (uint32_t)((qd >> 1) * (kBM * 128) + rowl * 128 + ((((qd & 1) * 4) ^ (rowl & 7)) << 4));
    const uint32_t prow_x = tc05::smem_u32(smem + kBOffP) + row_x, dsrow_x = tc05::smem_u32(smem + kBOffdS) + row_x;
    const uint32_t stg_x = tc05::smem_u32(smem + kBOffStg) + row_x;       // dQ staging: fp32, 16 columns = 4 chunks
    // bias of local key jl for query i: relc[(n_qpad - 1 - i) + jl]; the parity of that offset is fixed per thread
    const float* relc_t = s_relc + ((n_qpad - 1 - rowl) & 1) * cs - ((n_qpad - 1 - rowl) & 1) + jl0;

    // The tiles of all items form ONE sequence g = 0, 1, ...: after the math of tile g the thread stores P/dS(g), stages
    // dQ(g-1) and, when tile g-1 closed an item, drains that item's dK / dV — so the wait for the previous tile's GEMMs
    // and the drain sit behind a tile of math instead of forming a serial tail per item (37 % of an item in
    // profiles/r02_call3_attn_phase_trace.txt).
    int g = 0;
    int h_cur = -1, kt_cur = -1;
    int w = w0;
    Item cur = item_of(w0), prv = cur;             // the item of tile g, and of tile g - 1
    int it = first_tile(cur.kt);
    h_cur = cur.h; kt_cur = cur.kt;
    // per-row statistics of a query tile, fetched one tile ahead so the global-load latency hides behind the math
    auto load_stats = [&](bool valid, const Item& x, int it_, float& l_out, float& d_out) {
      l_out = -INFINITY; d_out = 0.f;
      if (valid) {
        const int in = it_ * kBM + rowl;
        if (in < p.Sq) {
          const long long ri = ((long long)x.b * p.H + x.h) * p.Sq + in;
          l_out = p.lse[ri];
          d_out = p.delta[ri];
        }
      }
    };
    // dQ of tile gp (complete in TMEM) -> fp32 staging tile in smem (the issuer reduces it into dq_accum)
    auto stage_dq = [&](int gp) {
      if (gp > 0) tc05::mbar_wait(bar_stg, (gp - 1) & 1);         // the reduce of the tile before has read the buffer
      uint32_t r[16];
      tc05::tmem_ld_32x16(tmem_row + 384 + (gp & 1) * 64 + qd * 16, r);
      tc05::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     :: "r"(stg_x ^ (uint32_t)(q << 4)), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                     : "memory");
    };
    // dV / dK rows of a finished item (key j0 + rowl), columns [16 qd, +16): TMEM -> bf16 -> global
    auto drain_dkv = [&](const Item& x) {
      const int h = x.h, b = x.b;
      const int j = x.kt * kBN + rowl;
      uint32_t rv[16], rk[16];
      tc05::tmem_ld_32x16(tmem_row + 256 + qd * 16, rv);
      tc05::tmem_ld_32x16(tmem_row + 320 + qd * 16, rk);
      tc05::tmem_ld_wait();
      if (j < p.Sk) {
        __nv_bfloat16* dvrow = p.dv + b * p.dv_stride_b + j * p.dv_stride_s + h * p.dv_stride_h + qd * 16;
        __nv_bfloat16* dkrow = p.dk + b * p.dk_stride_b + j * p.dk_stride_s + h * p.dk_stride_h + qd * 16;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 uv, uk;
          uv.x = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 0]), __uint_as_float(rv[c * 8 + 1]));
          uv.y = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 2]), __uint_as_float(rv[c * 8 + 3]));
          uv.z = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 4]), __uint_as_float(rv[c * 8 + 5]));
          uv.w = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 6]), __uint_as_float(rv[c * 8 + 7]));
          uk.x = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 0]), __uint_as_float(rk[c * 8 + 1]));
          uk.y = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 2]), __uint_as_float(rk[c * 8 + 3]));
          uk.z = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 4]), __uint_as_float(rk[c * 8 + 5]));
          uk.w = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 6]), __uint_as_float(rk[c * 8 + 7]));
          *reinterpret_cast<uint4*>(dvrow + c * 8) = uv;
          *reinterpret_cast<uint4*>(dkrow + c * 8) = uk;
        }
      }
    };
    float lse_nx, delta_nx;
    load_stats(true, cur, it, lse_nx, delta_nx);
    bool prev_closed_item = false;                 // tile g-1 was the last query tile of its item (`prv`)
    for (;; ++g) {
      const int h = cur.h, kt = cur.kt, b = cur.b;
      const int n = w - w0;
      const int j0 = kt * kBN;
      const int m_first = first_tile(kt);
      if (it == m_first) {
        // a new item.  Every compute thread is past the previous item's math (bias reads, bin updates): the bins can
        // be flushed and restaged if the (head, key tile) changed, and the key term of the item AFTER this one can
        // overwrite the buffer the previous item read.
        tc05::named_bar_sync(2, kBComputeThreads);
        if (HAS_REL && (h != h_cur || kt != kt_cur)) {
          flush_rel(h_cur, kt_cur);
          tc05::named_bar_sync(2, kBComputeThreads);
          stage_rel(h, kt);
          h_cur = h; kt_cur = kt;
          tc05::named_bar_sync(2, kBComputeThreads);
        }
        if (w + 1 < w1 && tid < kBN) {             // (read after the named barrier that opens item n + 1)
          const Item nx = next_item(cur);
          const int j = nx.kt * kBN + tid;
          s_kadd[((n + 1) & 1) * kBN + tid] =
              (j < p.Sk) ? (p.key_add ? p.key_add[(long long)nx.b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
        }
      }
      const float* kadd = s_kadd + (n & 1) * kBN + jl0;
      const bool cols_dead = j0 + jl0 >= p.Sk;          // all 32 keys of this thread are past the end (warp-uniform)
      const int i0 = it * kBM;
      const int i = i0 + rowl;
      const bool last_of_item = it + 1 == p.n_qt;
      // the tile after this one (for the prefetch of its row statistics)
      Item x_nx = cur;
      int it_nx = it + 1;
      bool has_nx = true;
      if (last_of_item) {
        x_nx = next_item(cur);
        it_nx = first_tile(x_nx.kt);
        has_nx = w + 1 < w1;
      }
      // +inf => p = exp2(s - inf) = 0 for dead rows and for rows whose softmax was empty (lse = -inf)
      const float lse2 = (lse_nx != -INFINITY) ? lse_nx * kLog2e : INFINITY;
      const float delta = delta_nx;
      load_stats(has_nx, x_nx, it_nx, lse_nx, delta_nx);
      const bool dead = cols_dead || i0 + quad * 32 >= p.Sq;      // warp-uniform
      const bool diag = CAUSAL && (j0 + kBN - 1 > i0);
      const bool tr = g < 3;
      if (tr) PVQA_TRACEB(2 + 6 * g);
      tc05::mbar_wait(bar_s, g & 1);
      tc05::tc_fence_after_sync();
      if (tr) PVQA_TRACEB(3 + 6 * g);

      uint32_t pw[16], dw[16];                 // bf16x2 words of this thread's P and dS columns
      if (dead) {
        tc05::tc_fence_before_sync();
        tc05::named_bar_arrive(1, kBSyncThreads);
#pragma unroll
        for (int x = 0; x < 16; ++x) { pw[x] = 0u; dw[x] = 0u; }
      } else {
          // ---- e = s * scale * log2e + bias - lse  (exp2 domain; 1/keep rides in the exponent with dropout);
          //      p = exp2(e);  ds = scale * p * (M dP / keep - delta):  with dropout pk = p / keep, pm = pk & M and
          //      ds = pm * (scale dP) + pk * (-scale keep_prob delta).  Eight columns at a time keeps registers short.
          const float lse_k = DROP ? lse2 - p.drop.m_shift : lse2;
          const float2 nl = make_float2(-lse_k, -lse_k);
          const float nd = -(DROP ? delta / p.drop.keep_scale : delta) * p.scale;
          const float2 nd2 = make_float2(nd, nd), sc2 = make_float2(p.scale, p.scale);
          const float4* ka4 = reinterpret_cast<const float4*>(kadd);
          const float2* rl2 = reinterpret_cast<const float2*>(relc_t + (n_qpad - 1 - rowl) - i0);
          const bool row_in = SCP && i < p.Sq && i >= p.scp_q0 && i < p.scp_q0 + p.scp_L;
          const uint8_t* scp_row = SCP ? p.scp_bucket + ((long long)b * p.scp_L + (i - p.scp_q0)) * p.scp_L : nullptr;
          // d_rel: diagonals lane (mod 32) of this warp's 32 x 32 block — or, when every offset of the block lies in one
          // of the two constant tails of the bias vector (T5 buckets: |j - i| >= 91 at 32 buckets / max distance 128),
          // the plain sum of the block credited to one offset of that tail (the caller bins offsets that share a value)
          float accp = 0.f, accn = 0.f;
          float2 far2 = make_float2(0.f, 0.f);
          const int d_lo = (j0 + jl0) - (i0 + quad * 32 + 31), d_hi = (j0 + jl0 + 31) - (i0 + quad * 32);
          const bool far = HAS_REL && p.rel_far > 0 && (d_lo >= p.rel_far || d_hi <= -p.rel_far);       // warp-uniform
          float2 s2[8], dp2[8];                   // 16 columns of S and of dP at a time
          uint32_t kw[8];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            if ((c8 & 1) == 0) {
              tc05::tmem_ld_32x16(tmem_row + jl0 + c8 * 8, *reinterpret_cast<uint32_t(*)[16]>(s2));
              tc05::tmem_ld_32x16(tmem_row + kBN + jl0 + c8 * 8, *reinterpret_cast<uint32_t(*)[16]>(dp2));
              if (DROP && c8 == 0) {               // the dropout bits (integer work) hide under the TMEM load
                const uint64_t ctr = rng_off + ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * p.drop.blk_per_row +
                                     (uint32_t)((j0 + jl0) >> 5);
                keep_shifted(keep_bits32(p.drop, ctr), kw);
              }
              tc05::tmem_ld_wait();
              if (c8 == 2) {
                // S/dP of this tile now live in registers: the issuer may overwrite TMEM with the next tile's
                tc05::tc_fence_before_sync();
                tc05::named_bar_arrive(1, kBSyncThreads);
              }
            }
            float2 e[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float4 ka = ka4[c8 * 2 + q];
              float2 b0 = make_float2(ka.x, ka.y), b1 = make_float2(ka.z, ka.w);
              if (HAS_REL) {
                b0 = add2(b0, rl2[c8 * 4 + 2 * q]);
                b1 = add2(b1, rl2[c8 * 4 + 2 * q + 1]);
              }
              e[2 * q] = add2(fma2(s2[(c8 & 1) * 4 + 2 * q], make_float2(p.sl2, p.sl2), b0), nl);
              e[2 * q + 1] = add2(fma2(s2[(c8 & 1) * 4 + 2 * q + 1], make_float2(p.sl2, p.sl2), b1), nl);
            }
            float* ef = reinterpret_cast<float*>(e);
            uint32_t bk0 = 0u, bk1 = 0u;          // SCP bucket ids of the 8 columns (8-column groups are in or out)
            bool scp_in = false;
            if (SCP) {
              const int jj = j0 + jl0 + c8 * 8 - p.scp_q0;
              scp_in = row_in && jj >= 0 && jj < p.scp_L;
              if (scp_in) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(scp_row + jj));
                bk0 = u.x; bk1 = u.y;
#pragma unroll
                for (int x = 0; x < 8; ++x) ef[x] += s_scp[((x < 4 ? bk0 : bk1) >> (8 * (x & 3))) & 31u];
              }
            }
            if (CAUSAL) {
              if (diag) {
#pragma unroll
                for (int x = 0; x < 8; ++x)
                  if (j0 + jl0 + c8 * 8 + x > i) ef[x] = -INFINITY;
              }
            }
            float dsv[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 pk = e[q];
              pk.x = fast_exp2(pk.x);
              pk.y = fast_exp2(pk.y);
              float2 ds;
              if (DROP) {
                float2 pm;
                pm.x = __uint_as_float(__float_as_uint(pk.x) & prmt(kw[7 - 2 * q], 0u, word_sel(c8)));
                pm.y = __uint_as_float(__float_as_uint(pk.y) & prmt(kw[6 - 2 * q], 0u, word_sel(c8)));
                ds = fma2(pm, mul2(dp2[(c8 & 1) * 4 + q], sc2), mul2(pk, nd2));
                pw[c8 * 4 + q] = f32x2_to_bf16x2(pm.x, pm.y);
              } else {
                ds = mul2(pk, fma2(dp2[(c8 & 1) * 4 + q], sc2, nd2));
                pw[c8 * 4 + q] = f32x2_to_bf16x2(pk.x, pk.y);
              }
              dw[c8 * 4 + q] = f32x2_to_bf16x2(ds.x, ds.y);
              dsv[2 * q] = ds.x; dsv[2 * q + 1] = ds.y;
            }
            if (SCP) {
              // d_scp[bucket] += dS on the OCR x OCR block: per-warp shared-memory bins (conflicting lanes serialise)
              if (p.d_scp && scp_in) {
#pragma unroll
                for (int x = 0; x < 8; ++x)
                  atomicAdd(s_dscp + warp * 32 + (((x < 4 ? bk0 : bk1) >> (8 * (x & 3))) & 31u), dsv[x]);
              }
            }
            if (HAS_REL) {
              if (p.d_rel) {
                if (far) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) far2 = add2(far2, make_float2(dsv[2 * q], dsv[2 * q + 1]));
                } else {
                  // this warp holds a 32x32 block (lane = row, register = column).  Lane L collects the diagonals
                  // d' = col - row == L (mod 32): one shuffle per column, two accumulators for the wrap.
#pragma unroll
                  for (int x = 0; x < 8; ++x) {
                    const int c = c8 * 8 + x;
                    const float vsh = __shfl_sync(0xffffffffu, dsv[x], c - lane);   // source lane taken mod 32
                    asm("{\n\t.reg .pred pq;\n\t"
                        "setp.le.s32 pq, %2, %3;\n\t"
                        "@pq add.f32 %0, %0, %4;\n\t"
                        "@!pq add.f32 %1, %1, %4;\n\t}"
                        : "+f"(accp), "+f"(accn) : "r"(lane), "r"(c), "f"(vsh));
                  }
                }
              }
            }
          }
          if (HAS_REL) {
            if (p.d_rel) {
              // window index x = jl - i + n_qpad - 1 with jl - il = (jl0 - 32 quad) + d'
              const int wpos = jl0 - quad * 32 + lane - i0 + n_qpad - 1;
              if (far) {
                float t = far2.x + far2.y;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) atomicAdd(s_drel + wpos, t);       // offset of the block's first (live) element
              } else {
                if (wpos >= 0 && wpos < n_win) atomicAdd(s_drel + wpos, accp);
                if (lane > 0 && wpos - 32 >= 0 && wpos - 32 < n_win) atomicAdd(s_drel + wpos - 32, accn);
              }
            }
          }
      }
      if (tr) PVQA_TRACEB(4 + 6 * g);
      // P/dS smem is still read by the GEMMs of tile g-1: wait for them right before the stores (long done by then)
      if (g > 0) {
        tc05::mbar_wait(bar_g, (g - 1) & 1);
        tc05::tc_fence_after_sync();
      }
      if (tr) PVQA_TRACEB(5 + 6 * g);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     :: "r"(prow_x ^ (uint32_t)(q << 4)), "r"(pw[4 * q]), "r"(pw[4 * q + 1]), "r"(pw[4 * q + 2]), "r"(pw[4 * q + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     :: "r"(dsrow_x ^ (uint32_t)(q << 4)), "r"(dw[4 * q]), "r"(dw[4 * q + 1]), "r"(dw[4 * q + 2]), "r"(dw[4 * q + 3])
                     : "memory");
      }
      if (tr) PVQA_TRACEB(6 + 6 * g);
      if (g > 0) {
        stage_dq(g - 1);                          // its GEMMs are complete (bar_g above)
        if (prev_closed_item) drain_dkv(prv);     // before this tile's GEMMs (accumulate = 0) overwrite dV / dK
      }
      tc05::fence_proxy_async_smem();
      tc05::tc_fence_before_sync();
      tc05::named_bar_arrive(3, kBSyncThreads);
      if (tr) PVQA_TRACEB(7 + 6 * g);
      prev_closed_item = last_of_item;
      prv = cur;
      // next tile of the sequence
      if (!last_of_item) {
        ++it;
      } else {
        if (n == 0) PVQA_TRACEB(28);
        if (n == 1) PVQA_TRACEB(29);
        if (++w >= w1) break;
        cur = x_nx;
        it = it_nx;
      }
    }
    // ---- tail of the sequence: dQ of the last tile, dV / dK of the last item, the bins of the last (head, key tile)
    ++g;                                          // (g now counts the tiles done)
    tc05::mbar_wait(bar_g, (g - 1) & 1);
    tc05::tc_fence_after_sync();
    stage_dq(g - 1);
    tc05::fence_proxy_async_smem();
    tc05::tc_fence_before_sync();
    tc05::named_bar_arrive(3, kBSyncThreads);
    drain_dkv(prv);
    if (HAS_REL) {
      tc05::named_bar_sync(2, kBComputeThreads);
      flush_rel(h_cur, kt_cur);
    }
  }
  PVQA_TRACEB(30);
  tc05::tc_fence_before_sync();
  __syncthreads();
  PVQA_TRACEB(31);
  if (is_issuer) tc05::tmem_dealloc(tmem_base, kBTmemCols);
}

}  // namespace pvqa
