// attn_fwd.cuh — K2 / K3 forward: persistent flash attention on tcgen05 (included by attn.cu).
//
// What the device taught (profiles/r02_call1_*, r02_call3_attn_phase_trace.txt):
//   * three structurally different round-1 forwards all ran at ~110 / 142 us on the bench shape because each spent 25+
//     issue slots per score on the CUDA cores — so this kernel is built around an instruction budget: packed fp32x2
//     adds / FMAs (new on sm_100), the relative bias read with 8-byte shared loads from two shifted copies, dropout
//     bit-sliced (eight logic ops decide 32 keys) and applied to the PACKED bf16x2 words through PRMT's sign-replicate
//     mode, Philox4x32-7, causal / SaL-SCP code only in the instantiations that need it: ~9 slots per score with
//     dropout, ~5.5 without, against 8 MUFU-pipe cycles per score-warp;
//   * the first cut of that kernel (two CTAs per SM, 8 softmax warps each, 64-key tiles) was no faster: with four
//     96-register warps per scheduler every dependent LDS / MUFU / barrier latency was exposed (31 % issue utilisation),
//     while the backward — 16 warps of ONE CTA per SM, 32 scores per thread, registers to spare for the compiler to
//     software-pipeline — reached ~80 % on the same kind of code.  The forward now has the backward's shape;
//   * `tcgen05.mma` issued under `if (lane == 0)` cost ~100 cycles each (R2UR waterfall); the issuer warp therefore
//     runs warp-uniformly and predicates only the instruction on tc05::elect_one().
// Structure:
//   * persistent, one CTA per SM walks a contiguous range of (head, query tile, batch) items; 16 softmax warps (TMEM
//     lane = query row, a quarter of a key tile's columns per thread) + 1 issuer warp (TMA + tcgen05.mma);
//   * the tiles of all items form ONE sequence: S_{g+1} = Q K^T is issued a tile ahead into the other TMEM buffer,
//     O += P_g V_g as soon as P_g is in smem, and the epilogue of an item (O / l -> bf16 rows) runs inside the first
//     tile of the next item, after its scores are in registers — no serial per-item tail;
//   * S is read from TMEM once and stays in registers; O accumulates in TMEM and is rescaled in place only when a row
//     maximum outgrows the running reference by more than 2^8 (lazy rescaling);
//   * key tiles are 128 wide; the remainder of a row is ONE tile of width 32, 64 or 128, masked past Sk (S = 327:
//     128 + 128 + 128).  Splitting the remainder into 64 + 32 saves padding but costs a tile's fixed overhead and
//     measured 10 % slower (see the launcher in attn.cu).
#pragma once

namespace pvqa {

constexpr int kFSoftmaxWarps = 16;
constexpr int kFSoftmaxThreads = kFSoftmaxWarps * 32;      // warp w: rows (w & 3) * 32 + lane, column quarter w >> 2
constexpr int kFThreads = kFSoftmaxThreads + 32;           // + the issuer warp (warp 16)
constexpr uint32_t kFTmemCols = 512;                       // S_g: [128 (g & 1), +128)   O: [256,320)
constexpr int kFStages = 3;                                // K / V rings: tile g in slot g % 3, loaded two tiles ahead
constexpr int kFTileBytes = kBN * kD * 2;                  // 16 KB
constexpr int kFOffQ = 0;                                  // 2 x 16 KB  Q of item n in buffer n & 1
constexpr int kFOffK = kFOffQ + 2 * kBM * kD * 2;          // 3 x 16 KB
constexpr int kFOffV = kFOffK + kFStages * kFTileBytes;    // 3 x 16 KB
constexpr int kFOffP = kFOffV + kFStages * kFTileBytes;    // 32 KB: P as two [128][64] K-major SW128 sub-tiles
constexpr int kFOffBar = kFOffP + kBM * kBN * 2;           // 160 KB
constexpr int kFOffXchg = kFOffBar + 192;                  // [2 tile parities][4 quarters][128 rows] + [4][128] floats
constexpr int kFOffFloats = kFOffXchg + 3 * 4 * kBM * 4;   // kadd[2][n_kpad], relc[2][cs], scp table[32]

// two shifted copies of the relative-bias window of a query tile: copy_k[a] = staged[a + k], staged[x] = bias of
// relative offset (x - 127) + (j0 - i0) ... see stage_rel.  cs % 32 == 16 puts the two copies 16 banks apart, so the
// 8-byte reads of a half-warp (even lanes -> one copy, odd lanes -> the other) never share a bank.
__host__ __device__ constexpr int f_rel_copy_stride(int n_kpad) { return ((n_kpad + 128 + 2 + 31) / 32) * 32 + 16; }

// ---- attention-probability dropout, shared by forward and backward ----------------------------------------------
// One Philox counter per (query row, block of 32 keys).  Two Philox4x32-7 calls give eight random words = eight bit
// planes: key e of the block draws the 8-bit uniform U_e = sum_k 2^k * bit_e(plane_k) and is kept iff U_e >= thr8, so
// the drop probability is thr8 / 256 (p quantised to 1/256; the 1/keep scale uses the quantised value).  The compare
// runs bit-sliced — eight 3-input logic ops decide all 32 keys — and leaves one keep BIT per key (bit e of the result).
// Masks are widened from bits with PRMT's sign-replicate mode (see keep_shifted): one instruction per bf16x2 pair.
// Everything thread-invariant (round keys, threshold planes) sits in the kernel parameters, i.e. in the constant bank.
struct AttnDrop {
  uint32_t rk[7][2];          // Philox round keys: seed + r * (W0, W1)
  uint32_t tmask[8];          // bit k of thr8 as an all-ones / all-zeros word
  float keep_scale;           // 1 / keep probability
  float m_shift;              // log2(keep_scale): the 1/keep factor rides in the softmax exponent
  uint32_t blk_per_row;       // ceil(Sk / 32)
  uint64_t offset;            // host Philox offset of this launch
  const unsigned long long* rng_base;    // optional device step counter (CUDA-graph replays), see common.cuh
};

// Philox4x32-7: the smallest round count that passes BigCrush (Salmon et al., SC'11); dropout masks need no more.
__device__ __forceinline__ uint4 philox4x32_7(uint4 ctr, const uint32_t (&rk)[7][2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint64_t p0 = (uint64_t)M0 * ctr.x, p1 = (uint64_t)M1 * ctr.z;
    ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ rk[r][0], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr.w ^ rk[r][1], (uint32_t)p0);
  }
  return ctr;
}
// keep bits of the 32-key block with Philox counter c: bit e set <=> key e of the block is kept
__device__ __forceinline__ uint32_t keep_bits32(const AttnDrop& d, uint64_t c) {
  const uint4 a = philox4x32_7(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5A17u, 0u), d.rk);
  const uint4 b = philox4x32_7(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5A18u, 0u), d.rk);
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t ge = 0xffffffffu;          // U >= T over the bits seen so far (LSB first): equal so far counts as >=
#pragma unroll
  for (int k = 0; k < 8; ++k) ge = (u[k] & ge) | (~d.tmask[k] & (u[k] | ge));     // t_k ? u & ge : u | ge
  return ge;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// w[j] = keep << j puts the flag of key 8*b + (7 - j) into the top bit of byte b.  PRMT with selector bit 3 set
// replicates a byte's top bit over the whole byte, so
//   bf16x2 mask of keys (8b + 2q, 8b + 2q + 1) = prmt(w[7 - 2q], w[6 - 2q], pair_sel(b))
//   fp32 mask of key 8b + r                     = prmt(w[7 - r], 0, word_sel(b))
__device__ __forceinline__ void keep_shifted(uint32_t keep, uint32_t (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = keep << j;
}
__host__ __device__ constexpr uint32_t pair_sel(int b) { return (uint32_t)(((12 + b) * 0x11) << 8 | ((8 + b) * 0x11)); }
__host__ __device__ constexpr uint32_t word_sel(int b) { return (uint32_t)((8 + b) * 0x1111); }

struct AttnFwdParams {
  __nv_bfloat16* o;
  float* lse;                 // (B,H,Sq)
  const float* rel_bias;      // (H, Sq+Sk-1) or null
  const float* key_add;       // (B, Sk) or null
  int B, H, Sq, Sk;
  long long o_stride_b, o_stride_s, o_stride_h;
  float sl2;                  // scale * log2(e): the softmax runs in the exp2 domain
  AttnDrop drop;
  // SaL spatial (SCP) bias: bias += scp_tab[h][scp_bucket[b][i-q0][j-q0]] on the OCR x OCR block
  const uint8_t* scp_bucket;  // (B, L, L) or null
  const float* scp_tab;       // (H, 32)
  int scp_q0, scp_L;
  // work decomposition
  int n_qt;                   // query tiles per (b, h)
  int n_kt;                   // key tiles: n_full of width 128, then one of width w_a and (if w_b) one of width w_b
  int n_full, w_a, w_b;       // remainder r = Sk % 128: r <= 32: (32,0); <= 64: (64,0); <= 96: (64,32); else (128,0)
  int n_items;                // H * n_qt * B
};

// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 — one issue slot for two lanes of work)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}

// named barrier over the four warps that share the 32 rows of a TMEM lane quadrant (ids 1..4, immediate operands:
// ptxas reserves five hardware barriers, not all sixteen)
__device__ __forceinline__ void quad_sync(int quad) {
  switch (quad) {
    case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
  }
}

#ifdef PVQA_ATTN_TRACE
#define PVQA_TRACEF(ev)                                                                             \
  do {                                                                                              \
    if (blockIdx.x < 64 && (ev) < 32) {                                                             \
      if (threadIdx.x == 0) g_attn_trace[blockIdx.x * 64 + (ev)] = clock64();                       \
      if (threadIdx.x == kFSoftmaxThreads) g_attn_trace[blockIdx.x * 64 + 32 + (ev)] = clock64();   \
    }                                                                                               \
  } while (0)
#else
#define PVQA_TRACEF(ev)
#endif


template <bool HAS_REL, bool DROP, bool CAUSAL, bool SCP>
__global__ void __launch_bounds__(kFThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an OFFSET into the __shared__ array so that the compiler
  // keeps the shared address space (32-bit LDS/STS instead of generic 64-bit LD/ST for every smem access)
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kFOffBar);   // [2] Q of item n landed, buffer n & 1     (TMA)
  uint64_t* bar_k = bar_q + 2;            // [3] K of tile g landed, slot g % 3                             (TMA)
  uint64_t* bar_v = bar_q + 5;            // [3] V of tile g landed                                         (TMA)
  uint64_t* bar_s = bar_q + 8;            // [2] S_g in TMEM buffer g & 1                                   (tcgen05.commit)
  uint64_t* bar_p = bar_q + 10;           // P_g in smem, O rescaled / read, S_g long in registers          (512 arrivals)
  uint64_t* bar_o = bar_q + 11;           // O (+)= P_g V_g done                                            (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 12);
  float* s_x = reinterpret_cast<float*>(smem + kFOffXchg);            // [2][4][128] row maxima, then [4][128] row sums
  const int n_kpad = p.n_kt * kBN;                                    // (upper bound of the columns the tiles cover)
  const int cs = f_rel_copy_stride(n_kpad);
  float* s_kadd = reinterpret_cast<float*>(smem + kFOffFloats);       // [2][n_kpad], -inf beyond Sk
  float* s_relc = s_kadd + 2 * n_kpad;                                 // [2][cs]
  float* s_scp = s_relc + (HAS_REL ? 2 * cs : 0);                      // [32] SCP table of the current head
  const int n_rel = p.Sq + p.Sk - 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == kFSoftmaxWarps;
  // this CTA's contiguous range of items; item w = (h * n_qt + qt) * B + b
  const int w0 = (int)((long long)blockIdx.x * p.n_items / gridDim.x);
  const int w1 = (int)((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
  auto decode = [&](int w, int& h, int& qt, int& b) {
    b = w % p.B;
    const int hq = w / p.B;
    qt = hq % p.n_qt;
    h = hq / p.n_qt;
  };
  // items are walked incrementally (b fastest, then query tile, then head): no run-time divisions after the first decode
  struct Item { int h, qt, b; };
  auto item_of = [&](int w) { Item x; decode(w, x.h, x.qt, x.b); return x; };
  auto next_item = [&](Item x) {
    if (++x.b == p.B) { x.b = 0; if (++x.qt == p.n_qt) { x.qt = 0; ++x.h; } }
    return x;
  };
  // key tile t: first key and width
  auto tile_start = [&](int t) { return t <= p.n_full ? t * kBN : p.n_full * kBN + p.w_a; };
  auto tile_width = [&](int t) { return t < p.n_full ? kBN : (t == p.n_full ? p.w_a : p.w_b); };
  auto tiles_of = [&](int qt) {       // key tiles an item visits (causal: up to the tile that holds the last row's key)
    if (!CAUSAL) return p.n_kt;
    const int i_last = min(qt * kBM + kBM - 1, p.Sq - 1);
    int nt = p.n_kt;
    while (nt > 1 && tile_start(nt - 1) > i_last) --nt;
    return nt;
  };
  PVQA_TRACEF(0);

  if (is_issuer) {
    if (lane == 0) {
      tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV);
      for (int x = 0; x < 2; ++x) tc05::mbar_init(bar_q + x, 1);
      for (int x = 0; x < kFStages; ++x) { tc05::mbar_init(bar_k + x, 1); tc05::mbar_init(bar_v + x, 1); }
      for (int x = 0; x < 2; ++x) tc05::mbar_init(bar_s + x, 1);
      tc05::mbar_init(bar_p, kFSoftmaxThreads); tc05::mbar_init(bar_o, 1);
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_slot, kFTmemCols);
    tc05::tmem_relinquish();
  }

  // ---- staging (softmax threads): additive vectors pre-multiplied by log2(e), the softmax runs in the exp2 domain ----
  // staged[x] = bias of relative index r = (Sq - 128 - i0) + x, x in [0, n_kpad + 128); row `rowl` of the tile reads
  // key j at x = j + 127 - rowl.  Two copies: copy_k[a] = staged[a + k].
  auto stage_rel = [&](int h, int i0) {
    const int base_r = p.Sq - kBM - i0;
    for (int x = tid; x < 2 * cs; x += kFSoftmaxThreads) {
      const int k = x >= cs ? 1 : 0;
      const int y = x - k * cs + k;
      const int r = base_r + y;
      s_relc[x] = (y < n_kpad + kBM && r >= 0 && r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
    }
    if (SCP && tid < 32) s_scp[tid] = p.scp_tab[h * 32 + tid] * kLog2e;
  };
  auto stage_kadd = [&](int buf, int b) {
    for (int j = tid; j < n_kpad; j += kFSoftmaxThreads)
      s_kadd[buf * n_kpad + j] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
  };
  if (!is_issuer && w0 < w1) {
    int h, qt, b;
    decode(w0, h, qt, b);
    stage_kadd(0, b);
    if (HAS_REL) stage_rel(h, qt * kBM);
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  PVQA_TRACEF(1);

  if (is_issuer) {
    // ------------------------------------------------------------------ issuer warp: warp-uniform control flow, one
    // elected lane per TMA / tcgen05 instruction (descriptors stay in uniform registers, see tc05::elect_one)
    if (w0 < w1) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      struct Cursor { int w, t, nt, h, qt, b; };
      auto enter = [&](Cursor& c, int w, Item x) { c.w = w; c.t = 0; c.h = x.h; c.qt = x.qt; c.b = x.b; c.nt = tiles_of(x.qt); };
      auto advance = [&](Cursor& c) {                      // to the next tile of this CTA's sequence; false at its end
        if (++c.t < c.nt) return true;
        if (c.w + 1 >= w1) return false;
        Item x; x.h = c.h; x.qt = c.qt; x.b = c.b;
        enter(c, c.w + 1, next_item(x));
        return true;
      };
      const uint32_t idesc_pv = tc05::idesc_bf16(kBM, kD, 0, 1);      // B = V is MN-major (d contiguous)
      const uint32_t smem0 = tc05::smem_u32(smem);
      const uint64_t p_desc = tc05::desc_sw128_k(smem0 + kFOffP);
      // The loads of a tile go out in two parts, each at the first point where the buffer it overwrites is known free
      // WITHOUT a wait of its own (every mbarrier wait costs ~100 cycles even when it succeeds at once):
      //   K_{g+2} -> slot of K_{g-1}: S_{g-1} completed before the softmax threads arrived on bar_p(g-1) (seen last turn)
      //   V_{g+2} -> slot of V_{g-1}, Q of a new item -> buffer of the item before the previous one: the softmax threads
      //   arrive on bar_p(g) only after waiting for P V of tile g-1 and for S_g, the last reader of that Q buffer.
      auto load_k = [&](const Cursor& c, int gx) {
        const int slot = gx % kFStages;
        const int j0 = tile_start(c.t);
        if (tc05::elect_one()) {
          tc05::mbar_expect_tx(bar_k + slot, kFTileBytes);
          tc05::tma_load_4d(smem + kFOffK + slot * kFTileBytes, &tmK, bar_k + slot, 0, c.h, j0, c.b);
        }
      };
      auto load_vq = [&](const Cursor& c, int gx) {
        const int slot = gx % kFStages;
        const int nq = (c.w - w0) & 1;
        const int j0 = tile_start(c.t);
        if (tc05::elect_one()) {
          if (c.t == 0) {
            tc05::mbar_expect_tx(bar_q + nq, kBM * kD * 2);
            tc05::tma_load_4d(smem + kFOffQ + nq * (kBM * kD * 2), &tmQ, bar_q + nq, 0, c.h, c.qt * kBM, c.b);
          }
          tc05::mbar_expect_tx(bar_v + slot, kFTileBytes);
          tc05::tma_load_4d(smem + kFOffV + slot * kFTileBytes, &tmV, bar_v + slot, 0, c.h, j0, c.b);
        }
      };
      // S of the tile under cursor c into TMEM buffer gx & 1 (operands must have landed)
      auto issue_s = [&](const Cursor& c, int gx) {
        const uint32_t idesc_qk = tc05::idesc_bf16(kBM, tile_width(c.t), 0, 0);
        const uint64_t qd = tc05::desc_sw128_k(smem0 + kFOffQ + ((c.w - w0) & 1) * (kBM * kD * 2));
        const uint64_t kd = tc05::desc_sw128_k(smem0 + kFOffK + (gx % kFStages) * kFTileBytes);
        if (tc05::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)      // 32 bytes per k-step inside the 128-byte swizzled row
            tc05::mma_bf16_ss(tmem_u + (gx & 1) * kBN, tc05::desc_step(qd, ks * 32), tc05::desc_step(kd, ks * 32),
                              idesc_qk, ks > 0);
          tc05::mma_commit(bar_s + (gx & 1));
        }
      };
      Cursor c_pv, c_s, c_ld;                    // tiles g, g + 1, g + 2
      enter(c_pv, w0, item_of(w0));
      c_s = c_pv; c_ld = c_pv;
      bool has_s, has_ld;
      load_k(c_ld, 0); load_vq(c_ld, 0);
      has_ld = advance(c_ld);
      if (has_ld) { load_k(c_ld, 1); load_vq(c_ld, 1); }
      has_s = has_ld;
      if (has_ld) has_ld = advance(c_ld);
      tc05::mbar_wait(bar_q, 0);
      tc05::mbar_wait(bar_k, 0);
      tc05::tc_fence_after_sync();
      issue_s(c_s, 0);
      if (has_s) advance(c_s);
      for (int g = 0;; ++g) {
        if (g < 3) PVQA_TRACEF(2 + 6 * g);
        if (has_ld) load_k(c_ld, g + 2);
        // S_{g+1}: operands landed; its TMEM buffer held S_{g-1}, which every softmax thread had in registers before
        // it arrived on bar_p(g-1)
        if (has_s) {
          tc05::mbar_wait(bar_k + ((g + 1) % kFStages), ((g + 1) / kFStages) & 1);
          if (c_s.t == 0) tc05::mbar_wait(bar_q + ((c_s.w - w0) & 1), ((c_s.w - w0) >> 1) & 1);
          tc05::tc_fence_after_sync();
          issue_s(c_s, g + 1);
          has_s = advance(c_s);
        }
        if (g < 3) PVQA_TRACEF(3 + 6 * g);
        // O (+)= P_g V_g: V_g landed, P_g written, O rescaled — or, for an item's first tile, the previous item's O
        // read out (its epilogue runs inside this tile, before the arrive on bar_p)
        tc05::mbar_wait(bar_v + (g % kFStages), (g / kFStages) & 1);
        if (g < 3) PVQA_TRACEF(4 + 6 * g);
        tc05::mbar_wait(bar_p, g & 1);
        tc05::tc_fence_after_sync();
        if (g < 3) PVQA_TRACEF(5 + 6 * g);
        const uint64_t vd = tc05::desc_sw128_k(smem0 + kFOffV + (g % kFStages) * kFTileBytes);
        const int ksteps = tile_width(c_pv.t) / 16;
        const uint32_t acc0 = c_pv.t > 0 ? 1u : 0u;
        if (tc05::elect_one()) {
          // A = P (K-major: sub-tile ks / 4, +32 B per step), B = V (MN-major, 16 keys = 2048 B per step)
          for (int ks = 0; ks < ksteps; ++ks)
            tc05::mma_bf16_ss(tmem_u + 2 * kBN, tc05::desc_step(p_desc, (ks >> 2) * (kBM * 128) + (ks & 3) * 32),
                              tc05::desc_step(vd, ks * 2048), idesc_pv, ks > 0 ? 1u : acc0);
          tc05::mma_commit(bar_o);
        }
        if (g < 3) PVQA_TRACEF(6 + 6 * g);
        if (has_ld) {
          load_vq(c_ld, g + 2);
          has_ld = advance(c_ld);
        }
        if (g < 3) PVQA_TRACEF(7 + 6 * g);
        if (!advance(c_pv)) break;
      }
    }
    __syncwarp();
  } else if (w0 < w1) {
    // ------------------------------------------------------------------ softmax warps: four threads per query row
    const int quad = warp & 3;                          // TMEM lane quadrant == 32-row group of the query tile
    const int qd = warp >> 2;                           // column quarter owned by this thread
    const int rowl = quad * 32 + lane;                  // row in the tile == TMEM lane
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint64_t rng_off = p.drop.offset + ((DROP && p.drop.rng_base) ? *p.drop.rng_base : 0ull);
    // this row's bias of key j: relc[j] (8-byte aligned for even j): copy (127 - rowl) & 1, element (127 - rowl) & ~1
    const float* relc = s_relc + ((127 - rowl) & 1) * cs + ((127 - rowl) & ~1);
    const uint32_t p_base = tc05::smem_u32(smem + kFOffP) + rowl * 128;
    float* lbuf = s_x + 2 * 4 * kBM;                    // [4][128] row sums of an item's epilogue

    // epilogue of a finished item (runs inside the next item's first tile, or after the last tile of the sequence):
    // combine the four quarter-row sums, O (TMEM) / l -> bf16 rows, lse (natural log)
    auto epilogue = [&](const Item& x, float m_fin, float l_fin) {
      const int h = x.h, qt = x.qt, b = x.b;
      const int i = qt * kBM + rowl;
      const bool rows_dead = qt * kBM + quad * 32 >= p.Sq;
      lbuf[qd * kBM + rowl] = l_fin;
      quad_sync(quad);
      const float l_tot = (lbuf[rowl] + lbuf[kBM + rowl]) + (lbuf[2 * kBM + rowl] + lbuf[3 * kBM + rowl]);
      quad_sync(quad);                                  // (the next epilogue writes the same slots)
      if (!rows_dead) {
        uint32_t r[16];
        tc05::tmem_ld_32x16(tmem_row + 2 * kBN + qd * 16, r);
        tc05::tmem_ld_wait();
        if (i < p.Sq) {
          const float inv = l_tot > 0.f ? (DROP ? p.drop.keep_scale : 1.f) / l_tot : 0.f;      // l carries 1/keep
          __nv_bfloat16* orow = p.o + (long long)b * p.o_stride_b + (long long)i * p.o_stride_s +
                                (long long)h * p.o_stride_h + qd * 16;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 u;
            u.x = f32x2_to_bf16x2(__uint_as_float(r[c * 8 + 0]) * inv, __uint_as_float(r[c * 8 + 1]) * inv);
            u.y = f32x2_to_bf16x2(__uint_as_float(r[c * 8 + 2]) * inv, __uint_as_float(r[c * 8 + 3]) * inv);
            u.z = f32x2_to_bf16x2(__uint_as_float(r[c * 8 + 4]) * inv, __uint_as_float(r[c * 8 + 5]) * inv);
            u.w = f32x2_to_bf16x2(__uint_as_float(r[c * 8 + 6]) * inv, __uint_as_float(r[c * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 8) = u;
          }
          if (p.lse && qd == 0)
            p.lse[((long long)b * p.H + h) * p.Sq + i] =
                l_tot > 0.f ? (m_fin + log2f(l_tot) - (DROP ? p.drop.m_shift : 0.f)) * (1.0f / kLog2e) : -INFINITY;
        }
      }
    };

    int g = 0;
    int h_cur = -1, qt_cur = -1;
    float m_run = -INFINITY, l_run = 0.f;
    Item cur = item_of(w0), prv = cur;
    for (int w = w0, n = 0; w < w1; ++w, ++n, prv = cur, cur = next_item(cur)) {
      const int h = cur.h, qt = cur.qt, b = cur.b;
      const int nt = tiles_of(qt);
      const int i0 = qt * kBM;
      const int i = i0 + rowl;
      const bool rows_dead = i0 + quad * 32 >= p.Sq;      // all 32 query rows of this warp are past the end
      if (n == 0) { h_cur = h; qt_cur = qt; }
      // every softmax thread is past the previous item's bias reads (they precede its last arrive on bar_p, and this
      // thread has since passed a wait that needed all of those arrivals): the buffers they read can be rewritten
      if (HAS_REL && (h != h_cur || qt != qt_cur)) {
        asm volatile("bar.sync 5, 512;" ::: "memory");
        stage_rel(h, i0);
        h_cur = h; qt_cur = qt;
        asm volatile("bar.sync 5, 512;" ::: "memory");
      }
      // key term of the NEXT item: the global loads go out now, the values are written to the other buffer right
      // before this item's last arrive on bar_p (readers of item n+1 are ordered behind it: arrive bar_p -> issuer ->
      // tcgen05.commit bar_o -> their wait on bar_o in item n+1's first tile)
      const bool stage_next = w + 1 < w1;
      float kv[2];
      int b_next = 0;
      if (stage_next) {
        b_next = next_item(cur).b;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = tid + u * kFSoftmaxThreads;
          kv[u] = (j < p.Sk) ? (p.key_add ? __ldg(p.key_add + (long long)b_next * p.Sk + j) * kLog2e : 0.f) : -INFINITY;
        }
      }
      const float* kadd = s_kadd + (n & 1) * n_kpad;
      // Philox counter of this row's first 32-key block (the block index is added per tile)
      const uint64_t drop_ctr = rng_off + ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * p.drop.blk_per_row;
      const uint8_t* scp_row = nullptr;                 // this row's bucket ids inside the OCR block, if it is in it
      if (SCP && i >= p.scp_q0 && i < p.scp_q0 + p.scp_L && i < p.Sq)
        scp_row = p.scp_bucket + ((long long)b * p.scp_L + (i - p.scp_q0)) * p.scp_L;
      const float m_prev = m_run, l_prev = l_run;       // the previous item's statistics, for its epilogue
      m_run = -INFINITY; l_run = 0.f;

      for (int t = 0; t < nt; ++t, ++g) {
        // A thread owns nc = width / 4 columns: 32, 16 or 8 (`nc` is CTA-uniform; the narrow cases run the same code
        // with the unused 8-column chunks switched off)
        const int nc = tile_width(t) >> 2;
        const int jb = tile_start(t) + qd * nc;            // first key of this thread's columns
        const bool tr = g < 3;
        if (tr) PVQA_TRACEF(2 + 6 * g);
        tc05::mbar_wait(bar_s + (g & 1), (g >> 1) & 1);
        tc05::tc_fence_after_sync();
        if (tr) PVQA_TRACEF(3 + 6 * g);
        float2 s[16];
        float* sf = reinterpret_cast<float*>(s);
        uint32_t kw[8];
        if (!rows_dead) {
          tc05::tmem_ld_32x32(tmem_row + (g & 1) * kBN + qd * nc, *reinterpret_cast<uint32_t(*)[32]>(s));
          if (DROP) {
            // the dropout bits of this thread's columns (ONE 32-key block, starting at bit jb & 31 of its keep word) are
            // pure integer work: done here it hides under the TMEM load and the latency-bound bias phase instead of
            // competing with the MUFU-bound exponential phase
            const uint32_t keep = keep_bits32(p.drop, drop_ctr + (uint32_t)(jb >> 5));
            keep_shifted(keep >> (jb & 31), kw);
          }
          tc05::tmem_ld_wait();
        }
        // Lazy rescaling: the running reference m_run moves only when a row's maximum exceeds it by more than 8 (a
        // factor 256 in the exp2 domain), so p = exp2(s - m_run) stays below 256 — exact enough in bf16, far from any
        // overflow in the fp32 row sum — and O in TMEM is touched only on those rare tiles, not on every new maximum.
        float m_new = m_run;
        bool grow = false;
        if (tr) PVQA_TRACEF(4 + 6 * g);
        if (!rows_dead) {
          // ---- biased scores in the exp2 domain and the max over this thread's columns ----
          const float4* ka4 = reinterpret_cast<const float4*>(kadd + jb);
          const float2* rl2 = reinterpret_cast<const float2*>(relc + jb);
          float mx = -INFINITY;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            if (c8 * 8 >= nc) break;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float4 ka = ka4[c8 * 2 + q];
              float2 b0 = make_float2(ka.x, ka.y), b1 = make_float2(ka.z, ka.w);
              if (HAS_REL) {
                b0 = add2(b0, rl2[c8 * 4 + 2 * q]);
                b1 = add2(b1, rl2[c8 * 4 + 2 * q + 1]);
              }
              s[c8 * 4 + 2 * q] = fma2(s[c8 * 4 + 2 * q], make_float2(p.sl2, p.sl2), b0);
              s[c8 * 4 + 2 * q + 1] = fma2(s[c8 * 4 + 2 * q + 1], make_float2(p.sl2, p.sl2), b1);
            }
            if (SCP) {
              const int jj = jb + c8 * 8 - p.scp_q0;                // block-relative; 8-column groups are in or out
              if (scp_row != nullptr && jj >= 0 && jj < p.scp_L) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(scp_row + jj));
#pragma unroll
                for (int k = 0; k < 8; ++k) sf[c8 * 8 + k] += s_scp[((k < 4 ? u.x : u.y) >> (8 * (k & 3))) & 31u];
              }
            }
            if (CAUSAL) {
              if (jb + c8 * 8 + 7 > i0) {                           // chunk may reach above the diagonal (warp-uniform)
#pragma unroll
                for (int x = 0; x < 8; ++x)
                  if (jb + c8 * 8 + x > i) sf[c8 * 8 + x] = -INFINITY;
              }
            }
#pragma unroll
            for (int x = 0; x < 8; x += 2) mx = fmaxf(fmaxf(mx, sf[c8 * 8 + x]), sf[c8 * 8 + x + 1]);
          }
          // ---- row max: exchange with the three threads that own the other quarters of this row (buffers alternate
          //      by tile parity, so the write of tile g+2 cannot overtake a partner's read of tile g) ----
          float* xbuf = s_x + (g & 1) * (4 * kBM);
          xbuf[qd * kBM + rowl] = mx;
          quad_sync(quad);
          mx = fmaxf(fmaxf(xbuf[rowl], xbuf[kBM + rowl]), fmaxf(xbuf[2 * kBM + rowl], xbuf[3 * kBM + rowl]));
          grow = mx > m_run + 8.f;                // also true for the first live tile (m_run = -inf)
          if (grow) m_new = mx;
        }
        if (tr) PVQA_TRACEF(5 + 6 * g);
        if (g > 0) {
          // O += P_{g-1} V_{g-1} has completed (it was issued a tile ago): the P buffer may be overwritten, and O may be
          // read out (first tile of an item: the previous item's epilogue) or rescaled
          tc05::mbar_wait(bar_o, (g - 1) & 1);
          tc05::tc_fence_after_sync();
        }
        if (t == 0) {
          if (n > 0) epilogue(prv, m_prev, l_prev);
        } else if (!rows_dead && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_run - m_new) : 1.f;          // exp2(-inf) = 0 for a first live tile
          uint32_t r[16];                                     // this thread's 16 of the 64 output columns
          tc05::tmem_ld_32x16(tmem_row + 2 * kBN + qd * 16, r);
          tc05::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 16; ++x) r[x] = __float_as_uint(__uint_as_float(r[x]) * alpha);
          tc05::tmem_st_32x16(tmem_row + 2 * kBN + qd * 16, r);
          tc05::tmem_st_wait();
          l_run *= alpha;
        }
        if (tr) PVQA_TRACEF(6 + 6 * g);
        if (!rows_dead) {
          // ---- p = exp2(s - m) (1/keep folded in), row sum, dropout on the packed words, bf16 P -> smem ----
          const float m_sub = ((m_new == -INFINITY) ? 0.f : m_new) - (DROP ? p.drop.m_shift : 0.f);
          const float2 nm = make_float2(-m_sub, -m_sub);
          float2 sum2 = make_float2(0.f, 0.f);
          // this thread's first 16-byte chunk (8 keys) of the P row, 128-byte swizzle folded in: chunk c8 of the thread
          // lives at pst ^ (c8 << 4)  (the thread's chunks never cross a 64-key sub-tile or carry in the chunk index)
          const int ch0 = (qd * nc) >> 3;
          const uint32_t pst = p_base + (uint32_t)((ch0 >> 3) * (kBM * 128)) + (uint32_t)((((ch0 & 7) ^ (rowl & 7))) << 4);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {                       // 8 keys = one 16-byte chunk of the P row
            if (c8 * 8 >= nc) break;
            uint32_t pw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 e = add2(s[c8 * 4 + q], nm);
              e.x = fast_exp2(e.x);
              e.y = fast_exp2(e.y);
              sum2 = add2(sum2, e);
              pw[q] = f32x2_to_bf16x2(e.x, e.y);
              if (DROP) pw[q] &= prmt(kw[7 - 2 * q], kw[6 - 2 * q], pair_sel(c8));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         :: "r"(pst ^ (uint32_t)(c8 << 4)), "r"(pw[0]), "r"(pw[1]), "r"(pw[2]), "r"(pw[3]) : "memory");
          }
          l_run += sum2.x + sum2.y;
          m_run = m_new;
        }
        if (stage_next && t + 1 == nt) {
          float* dst = s_kadd + ((n + 1) & 1) * n_kpad;
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (tid + u * kFSoftmaxThreads < n_kpad) dst[tid + u * kFSoftmaxThreads] = kv[u];
          for (int j = tid + 2 * kFSoftmaxThreads; j < n_kpad; j += kFSoftmaxThreads)     // Sk > 1024 only
            dst[j] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b_next * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
        }
        tc05::fence_proxy_async_smem();
        tc05::tc_fence_before_sync();
        tc05::mbar_arrive(bar_p);
        if (tr) PVQA_TRACEF(7 + 6 * g);
      }
      if (n == 0) PVQA_TRACEF(28);
      if (n == 1) PVQA_TRACEF(29);
    }
    // the last item's epilogue
    tc05::mbar_wait(bar_o, (g - 1) & 1);
    tc05::tc_fence_after_sync();
    epilogue(prv, m_run, l_run);
  }
  PVQA_TRACEF(30);
  tc05::tc_fence_before_sync();
  __syncthreads();
  PVQA_TRACEF(31);
  if (is_issuer) tc05::tmem_dealloc(tmem_base, kFTmemCols);
}

}  // namespace pvqa
