// attn_fwd3.cuh — third-generation attention forward (same contract as pvqa_attn_fwd; included by attn.cu).
//
// OPT-IN (PVQA_ATTN_FWD_V3=1), written after the first device run of v2 and NOT yet run on a device itself.
// What that run showed (profiles/r01_optin_kernels_timing.log): v2 is correct, executes half the instructions per
// score of attn_fwd_kernel, and is still 5 % slower — with one query row per thread a CTA has only four softmax warps
// (8 per SM instead of 16), and the kernel is bound by latency hiding, not by instruction count.  v3 therefore keeps
// the product kernel's thread mapping — 256 softmax threads, two threads per query row, each owning 64 of the 128
// columns, 16 softmax warps per SM — and takes from v2 only the mechanisms that remove work and stalls:
//   * S is read from TMEM once into 64 registers per thread and stays there across the row-max exchange (the
//     product kernel writes the biased scores back to TMEM and reads them again);
//   * the exchange is a 64-thread named barrier between the two warps that share rows, not a CTA barrier;
//   * O accumulates in TMEM across key tiles and is rescaled in place only when a row max of the warp grew (the
//     product kernel reads O_j back after every tile and merges it in registers);
//   * a separate issuer warp launches S_{t+1} = Q K_{t+1}^T as soon as all softmax threads hold S_t in registers, and
//     P_t V_t as soon as P_t is in shared memory, so the tensor pipe runs under the softmax instead of between its
//     phases (protocol: tools/attn_v2_protocol_sim.py, same barriers as v2 with 256 arrivals);
//   * the T5 bias is read with 128-bit shared loads from four shifted copies (pvqa_f3 layout, host-checked).
// Registers: 384 threads start at 80; the issuer warpgroup shrinks to 32, the two softmax warpgroups grow to 104.
// Dropout masks, score definition, lse and output are the definitions of attn_fwd_kernel.
#pragma once
#include "attn_fwd2_layout.h"

namespace pvqa {

constexpr int kF3SoftmaxThreads = 256;          // warps 0..7: row = (warp & 3) * 32 + lane, column half = warp >> 2
constexpr int kF3Threads = kF3SoftmaxThreads + 128;  // + the issuer warpgroup (warp 8 works, 9..11 only donate registers)
constexpr int kF3RegsSoftmax = 104;
constexpr int kF3RegsIssuer = 32;            // 256 * 104 + 128 * 32 == 384 * 80
constexpr uint32_t kF3TmemCols = 256;           // S: [0,128)   O: [128,192)
constexpr int kF3OffQ = 0;
constexpr int kF3OffKV = kF3OffQ + kBM * kD * 2;          // three 16 KB buffers: K_0 -> 0, K_t -> 2, V_t -> t odd ? 0 : 1
constexpr int kF3OffP = kF3OffKV + 3 * kKVBuf;            // 64 KB: P as two [128][64] K-major SW128 sub-tiles
constexpr int kF3OffBar = kF3OffP + kBM * kBN * 2;        // 96 KB
constexpr int kF3OffXchg = kF3OffBar + 128;               // [2 tile parities][2 halves][128 rows] floats
constexpr int kF3OffFloats = kF3OffXchg + 2 * 2 * kBM * 4;  // key_add[n_kpad], then 4 shifted copies of the bias vector

#ifdef PVQA_ATTN_TRACE
#define PVQA_TRACE3(ev)                                                                             \
  do {                                                                                              \
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z < 64 && (ev) < 32) {                       \
      if (threadIdx.x == 0) g_attn_trace[blockIdx.z * 64 + (ev)] = clock64();                       \
      if (threadIdx.x == kF3SoftmaxThreads) g_attn_trace[blockIdx.z * 64 + 32 + (ev)] = clock64();  \
    }                                                                                               \
  } while (0)
#else
#define PVQA_TRACE3(ev)
#endif

// SCP: the SaL spatial bias is compiled in only for launches that carry one (it costs registers in every variant
// that merely might)
// 64-thread named barrier between the two warps that own the two halves of the same 32 rows (ids 1..4, immediate
// operands so that ptxas reserves five barriers, not all sixteen)
__device__ __forceinline__ void pair_sync(int quad) {
  switch (quad) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <bool HAS_REL, bool DROP, bool SCP>
__global__ void __launch_bounds__(kF3Threads, 2)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kF3OffBar);
  uint64_t* bar_k = bar_q + 1;            // [2] by tile parity (TMA transaction barriers)
  uint64_t* bar_v = bar_q + 3;            // [2]
  uint64_t* bar_s = bar_q + 5;            // S_t in TMEM                 (tcgen05.commit)
  uint64_t* bar_sfree = bar_q + 6;        // S_t copied to registers     (256 arrivals)
  uint64_t* bar_p = bar_q + 7;            // P_t in smem, O rescaled     (256 arrivals)
  uint64_t* bar_o = bar_q + 8;            // O += P_t V_t done           (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 9);
  float* s_x = reinterpret_cast<float*>(smem + kF3OffXchg);            // [2][2][128]
  const int n_tiles_all = (p.Sk + kBN - 1) / kBN;
  const int n_kpad = n_tiles_all * kBN;
  const int cs = pvqa_f3::rel_copy_stride(p.Sq, n_kpad);
  float* s_kadd = reinterpret_cast<float*>(smem + kF3OffFloats);      // [n_kpad], -inf beyond Sk
  float* s_relc = s_kadd + n_kpad;                                     // [4][cs]
  float* s_scp = s_relc + (HAS_REL ? 4 * cs : 0);                      // [32] SCP table of this head (SaL)
  const int n_rel = p.Sq + p.Sk - 1;
  constexpr bool has_scp = SCP && HAS_REL;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * kBM;
  const int h = blockIdx.y, b = blockIdx.z;
  const bool is_issuer_wg = warp >= kF3SoftmaxThreads / 32;       // warpgroup-uniform
  const bool is_issuer_warp = warp == kF3SoftmaxThreads / 32;
  PVQA_TRACE3(0);

  int n_tiles = n_tiles_all;
  if (p.causal) n_tiles = min(n_tiles, min(i0 + kBM - 1, p.Sq - 1) / kBN + 1);

  if (is_issuer_warp) {
    if (lane == 0) {
      tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV);
      tc05::mbar_init(bar_q, 1);
      tc05::mbar_init(bar_k, 1); tc05::mbar_init(bar_k + 1, 1);
      tc05::mbar_init(bar_v, 1); tc05::mbar_init(bar_v + 1, 1);
      tc05::mbar_init(bar_s, 1); tc05::mbar_init(bar_sfree, kF3SoftmaxThreads);
      tc05::mbar_init(bar_p, kF3SoftmaxThreads); tc05::mbar_init(bar_o, 1);
      tc05::fence_barrier_init();
      // first loads go out before the bias staging; buffer 2 is free, so K_1 can be fetched at once as well
      tc05::mbar_expect_tx(bar_q, kBM * kD * 2);
      tc05::tma_load_4d(smem + kF3OffQ, &tmQ, bar_q, 0, h, i0, b);
      tc05::mbar_expect_tx(bar_k, kKVBuf);
      tc05::tma_load_4d(smem + kF3OffKV, &tmK, bar_k, 0, h, 0, b);
      tc05::mbar_expect_tx(bar_v, kKVBuf);
      tc05::tma_load_4d(smem + kF3OffKV + kKVBuf, &tmV, bar_v, 0, h, 0, b);
      if (n_tiles > 1) {
        tc05::mbar_expect_tx(bar_k + 1, kKVBuf);
        tc05::tma_load_4d(smem + kF3OffKV + 2 * kKVBuf, &tmK, bar_k + 1, 0, h, kBN, b);
      }
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_slot, kF3TmemCols);
    tc05::tmem_relinquish();
  } else if (!is_issuer_wg) {
    // stage the additive vectors, pre-multiplied by log2(e): the softmax runs in the exp2 domain
    for (int j0s = tid; j0s < n_kpad; j0s += 4 * kF3SoftmaxThreads) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0s + u * kF3SoftmaxThreads;
        v[u] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0s + u * kF3SoftmaxThreads < n_kpad) s_kadd[j0s + u * kF3SoftmaxThreads] = v[u];
    }
    if (HAS_REL) {
      // copy_k[x] = staged[x + k]; staged[y] = rel_bias[h][y - pad] * log2e inside [pad, pad + n_rel), 0 outside
      const int n = 4 * cs;
      for (int x0 = tid; x0 < n; x0 += 4 * kF3SoftmaxThreads) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = x0 + u * kF3SoftmaxThreads;
          const int r = pvqa_f3::rel_copy_source(idx, cs, p.Sq);
          v[u] = (idx < n && r >= 0 && r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (x0 + u * kF3SoftmaxThreads < n) s_relc[x0 + u * kF3SoftmaxThreads] = v[u];
      }
      if (has_scp && tid < 32) s_scp[tid] = p.scp_tab[h * 32 + tid] * kLog2e;
    }
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  PVQA_TRACE3(1);

  if (is_issuer_wg) {
    // ------------------------------------------------------------------ issuer: TMA + tcgen05.mma, one thread
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kF3RegsIssuer));
    if (is_issuer_warp && lane == 0) {
      const uint32_t idesc_qk = tc05::idesc_bf16(kBM, kBN, 0, 0);
      const uint32_t idesc_pv = tc05::idesc_bf16(kBM, kD, 0, 1);      // B = V is MN-major (d contiguous)
      const uint32_t q_addr = tc05::smem_u32(smem + kF3OffQ), kv_addr = tc05::smem_u32(smem + kF3OffKV);
      const uint32_t p_addr = tc05::smem_u32(smem + kF3OffP);
      // Buffer plan (three 16 KB buffers): K_0 -> 0, K_t (t >= 1) -> 2, V_t -> (t odd ? 0 : 1).
      //   buffer 2 is free as soon as S_t = Q K_t^T has completed, so K_{t+1} is fetched a whole tile before it is
      //   needed (the product kernel and v2 fetch it behind P_{t-1} V_{t-1}, which puts the TMA latency on the path
      //   to S_{t+1}); V_{t+1} takes the buffer of V_{t-1} (or of K_0) and is not needed before the end of tile t+1.
      auto k_buf = [&](int t) { return kv_addr + (t == 0 ? 0 : 2) * kKVBuf; };
      auto v_buf = [&](int t) { return kv_addr + ((t & 1) ? 0 : 1) * kKVBuf; };
      tc05::mbar_wait(bar_q, 0);
      tc05::mbar_wait(bar_k, 0);
      tc05::tc_fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks)
        tc05::mma_bf16_ss(tmem_base, tc05::smem_desc_sw128(q_addr + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(k_buf(0) + ks * 32, 16, 1024), idesc_qk, ks > 0);
      tc05::mma_commit(bar_s);
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = t * kBN;
        if (t + 1 < n_tiles) {
          // S_{t+1}: K_{t+1} landed (fetched one tile ago) and the softmax warps hold S_t in registers
          tc05::mbar_wait(bar_k + ((t + 1) & 1), ((t + 1) >> 1) & 1);
          tc05::mbar_wait(bar_sfree, t & 1);
          tc05::tc_fence_after_sync();
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)
            tc05::mma_bf16_ss(tmem_base, tc05::smem_desc_sw128(q_addr + ks * 32, 16, 1024),
                              tc05::smem_desc_sw128(k_buf(t + 1) + ks * 32, 16, 1024), idesc_qk, ks > 0);
          tc05::mma_commit(bar_s);
          PVQA_TRACE3(2 + 5 * t);                 // issuer: S_{t+1} issued
          if (t + 2 < n_tiles) {
            // K_{t+2} reuses buffer 2 as soon as S_{t+1} has been computed from it (a few hundred cycles from now)
            tc05::mbar_wait(bar_s, (t + 1) & 1);
            tc05::mbar_expect_tx(bar_k + (t & 1), kKVBuf);
            tc05::tma_load_4d(smem + kF3OffKV + 2 * kKVBuf, &tmK, bar_k + (t & 1), 0, h, j0 + 2 * kBN, b);
          }
          // V_{t+1} takes V_{t-1}'s buffer (O += P_{t-1} V_{t-1} must have completed) or, for t = 0, K_0's
          if (t >= 1) tc05::mbar_wait(bar_o, (t - 1) & 1);
          tc05::mbar_expect_tx(bar_v + ((t + 1) & 1), kKVBuf);
          tc05::tma_load_4d(smem + kF3OffKV + (((t + 1) & 1) ? 0 : 1) * kKVBuf, &tmV, bar_v + ((t + 1) & 1), 0, h,
                            j0 + kBN, b);
        }
        // O (+)= P_t V_t: V_t landed, P_t written (and O rescaled) by the softmax warps
        tc05::mbar_wait(bar_v + (t & 1), (t >> 1) & 1);
        tc05::mbar_wait(bar_p, t & 1);
        tc05::tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ++ks)
          tc05::mma_bf16_ss(tmem_base + kBN,
                            tc05::smem_desc_sw128(p_addr + (ks >> 2) * (kBM * 128) + (ks & 3) * 32, 16, 1024),
                            tc05::smem_desc_sw128(v_buf(t) + ks * 2048, 16, 1024), idesc_pv, (t > 0 || ks > 0) ? 1u : 0u);
        tc05::mma_commit(bar_o);
        PVQA_TRACE3(3 + 5 * t);                   // issuer: PV_t issued
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps: two threads per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kF3RegsSoftmax));
    const int rowl = (warp & 3) * 32 + lane;            // row in the tile == TMEM lane
    const int half = warp >> 2;                         // column half owned by this thread
    const int quad = warp & 3;                          // warps w and w + 4 own the two halves of the same rows
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int i = i0 + rowl;
    const bool rows_dead = i0 + (warp & 3) * 32 >= p.Sq;      // all 32 query rows of this warp are past the end
    const float sl2 = p.scale * kLog2e;
    const float* relc = s_relc + pvqa_f3::rel_copy_row_base(p.Sq, i, cs);    // bias of key j for this row: relc[j]
    const float m_shift = DROP ? log2f(p.drop_scale) : 0.f;     // 1/keep folded into the exponent (see attn_fwd_kernel)
    const uint32_t thr4 = p.drop_thr8 * 0x01010101u;
    const uint64_t drop_row = ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * (uint64_t)((p.Sk + 15) >> 4);
    const uint64_t rng_off = p.offset + ((DROP && p.rng_base) ? *p.rng_base : 0ull);
    const uint8_t* scp_row = nullptr;
    if (has_scp && i >= p.scp_q0 && i < p.scp_q0 + p.scp_L && i < p.Sq)
      scp_row = p.scp_bucket + ((long long)b * p.scp_L + (i - p.scp_q0)) * p.scp_L;
    uint8_t* prow = smem + kF3OffP + half * (kBM * 128) + rowl * 128;

    float m_run = -INFINITY, l_run = 0.f;
    float sf[64];                                       // this thread's half of the score row, updated in place

    for (int t = 0; t < n_tiles; ++t) {
      const int j0 = t * kBN;
      const int jh = j0 + half * 64;                    // first key of this thread's half
      const bool diag = p.causal && (j0 + kBN - 1 > i0);     // tile touches the diagonal (CTA-uniform)
      tc05::mbar_wait(bar_s, t & 1);
      tc05::tc_fence_after_sync();
      PVQA_TRACE3(2 + 5 * t);                     // softmax: S_t ready
      // ---- S_t -> registers (once).  The second 32-column chunk is in flight while the first one gets its bias;
      //      once both have landed the TMEM buffer goes back to the issuer ----
      const bool live0 = jh < p.Sk && !rows_dead, live1 = jh + 32 < p.Sk && !rows_dead;      // warp-uniform
      if (live0) {
        tc05::tmem_ld_32x32(tmem_row + half * 64, reinterpret_cast<uint32_t(&)[32]>(sf[0]));
        tc05::tmem_ld_wait();
      }
      if (live1) tc05::tmem_ld_32x32(tmem_row + half * 64 + 32, reinterpret_cast<uint32_t(&)[32]>(sf[32]));

      // ---- biased scores in the exp2 domain and the max over this thread's 64 columns ----
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int jb = jh + c * 32;
        if (c == 1) {
          tc05::tmem_ld_wait();
          tc05::tc_fence_before_sync();
          tc05::mbar_arrive(bar_sfree);
          PVQA_TRACE3(3 + 5 * t);                 // softmax: S_t in registers
        }
        if (!(c == 0 ? live0 : live1)) continue;
        const float4* ka4 = reinterpret_cast<const float4*>(s_kadd + jb);
        const float4* rl4 = reinterpret_cast<const float4*>(relc + jb);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 bs = ka4[q];
          if (HAS_REL) {
            const float4 rl = rl4[q];
            bs.x += rl.x; bs.y += rl.y; bs.z += rl.z; bs.w += rl.w;
          }
          const int x = c * 32 + q * 4;
          sf[x + 0] = fmaf(sf[x + 0], sl2, bs.x);
          sf[x + 1] = fmaf(sf[x + 1], sl2, bs.y);
          sf[x + 2] = fmaf(sf[x + 2], sl2, bs.z);
          sf[x + 3] = fmaf(sf[x + 3], sl2, bs.w);
        }
        if (has_scp && scp_row != nullptr && jb + 32 > p.scp_q0 && jb < p.scp_q0 + p.scp_L) {
          float sb[32];
          load_scp32(scp_row, jb - p.scp_q0, p.scp_L, s_scp, sb);
#pragma unroll
          for (int x = 0; x < 32; ++x) sf[c * 32 + x] += sb[x];
        }
        if (diag) {
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (jb + x > i) sf[c * 32 + x] = -INFINITY;
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) mx = fmaxf(mx, sf[c * 32 + x]);
      }
      // ---- row max: exchange with the thread that owns the other half of this row (buffers alternate by tile
      //      parity, so the write of tile t+2 cannot overtake the partner's read of tile t: barrier t+1 lies between)
      float* xbuf = s_x + (t & 1) * (2 * kBM);
      xbuf[half * kBM + rowl] = mx;
      pair_sync(quad);
      mx = fmaxf(mx, xbuf[(half ^ 1) * kBM + rowl]);
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_safe);      // 1 when the max did not move, 0 for the first live tile
      const float m_sub = m_safe - m_shift;
      PVQA_TRACE3(4 + 5 * t);                     // softmax: bias + max + exchange done

      if (t > 0) {
        // O += P_{t-1} V_{t-1} has completed: O may be rescaled and the P buffer may be overwritten
        tc05::mbar_wait(bar_o, (t - 1) & 1);
        tc05::tc_fence_after_sync();
        if (!rows_dead && __any_sync(0xffffffffu, m_new > m_run)) {
          uint32_t r[32];                                   // this thread's 32 of the 64 output columns
          tc05::tmem_ld_32x32(tmem_row + kBN + half * 32, r);
          tc05::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; ++x) r[x] = __float_as_uint(__uint_as_float(r[x]) * alpha);
          tc05::tmem_st_32x32(tmem_row + kBN + half * 32, r);
          tc05::tmem_st_wait();
        }
      }
      PVQA_TRACE3(5 + 5 * t);                     // softmax: O_{t-1} complete (and rescaled)

      // ---- p = exp2(s - m) (dropout folded in), row sum, bf16 P -> smem (K-major, 128B swizzle), sub-tile = half ----
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (jh + c * 32 >= p.Sk || rows_dead) {            // dead chunk: P must still be zero for the PV MMA
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(prow + (((c * 4 + q) ^ (rowl & 7)) * 16)) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        float* pv = sf + c * 32;                            // in place: the scores are not needed afterwards
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          pv[x] = fast_exp2(pv[x] - m_sub);
          sum += pv[x];
        }
        if (DROP) {
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            const uint4 kb = keep_bytes16(p.seed, rng_off, drop_row + ((jh + c * 32) >> 4) + g2, thr4);
            const uint32_t kw[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
#pragma unroll
              for (int bb = 0; bb < 4; ++bb) {
                const int x = g2 * 16 + w * 4 + bb;
                pv[x] = __uint_as_float(__float_as_uint(pv[x]) & PVQA_BYTE_MASK(kw[w], bb));
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c * 4 + q) ^ (rowl & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = pack8(pv + q * 8);
        }
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      tc05::fence_proxy_async_smem();
      tc05::tc_fence_before_sync();
      tc05::mbar_arrive(bar_p);
      PVQA_TRACE3(6 + 5 * t);                     // softmax: P_t stored
    }

    // ---- epilogue: combine the two half-row sums, O (TMEM) / l -> bf16 rows, lse (natural log) ----
    float* xbuf = s_x + (n_tiles & 1) * (2 * kBM);      // the parity the last tile did not use
    xbuf[half * kBM + rowl] = l_run;
    pair_sync(quad);
    const float l_tot = l_run + xbuf[(half ^ 1) * kBM + rowl];       // carries the 1/keep factor when DROP
    tc05::mbar_wait(bar_o, (n_tiles - 1) & 1);
    tc05::tc_fence_after_sync();
    PVQA_TRACE3(29);                              // softmax: last product done
    if (!rows_dead) {
      const float inv = l_tot > 0.f ? (DROP ? p.drop_scale : 1.f) / l_tot : 0.f;
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + kBN + half * 32, r);
      tc05::tmem_ld_wait();
      if (i < p.Sq) {
        float ov[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) ov[x] = __uint_as_float(r[x]) * inv;
        __nv_bfloat16* orow = p.o + (long long)b * p.o_stride_b + (long long)i * p.o_stride_s +
                              (long long)h * p.o_stride_h + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = pack8(ov + c * 8);
        if (p.lse && half == 0)
          p.lse[((long long)b * p.H + h) * p.Sq + i] =
              l_tot > 0.f ? (m_run + log2f(l_tot) - m_shift) * (1.0f / kLog2e) : -INFINITY;
      }
    }
  }
  PVQA_TRACE3(30);
  tc05::tc_fence_before_sync();
  __syncthreads();
  PVQA_TRACE3(31);
  if (is_issuer_warp) tc05::tmem_dealloc(tmem_base, kF3TmemCols);
}

}  // namespace pvqa

extern "C" int pvqa_attn_fwd_v3(const void* q, const void* k, const void* v, void* o, float* lse,
                                const float* rel_bias, const float* key_add, int64_t B, int64_t H, int64_t Sq,
                                int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                                int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t v_stride_b,
                                int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b, int64_t o_stride_s,
                                int64_t o_stride_h, float scale, int causal, float dropout_p, uint64_t seed,
                                uint64_t offset, const uint8_t* scp_bucket, const float* scp_table, int64_t scp_q0,
                                int64_t scp_L, void* stream) {
  using namespace pvqa;
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "attn_fwd_v3: dropout_p must be in [0,1)");
  if (scp_bucket) {
    PVQA_REQUIRE(rel_bias && scp_table, PVQA_ERR_NULL, "attn_fwd_v3: the SCP bias needs rel_bias and scp_table");
    PVQA_REQUIRE(!causal && Sq == Sk && scp_q0 >= 0 && scp_L > 0 && scp_q0 + scp_L <= Sk, PVQA_ERR_SHAPE,
                 "attn_fwd_v3: bad SCP block");
    PVQA_REQUIRE(scp_q0 % 16 == 0 && scp_L % 16 == 0 && aligned16(scp_bucket), PVQA_ERR_ALIGN,
                 "attn_fwd_v3: SCP block offset/size must be multiples of 16");
  }
  PVQA_REQUIRE(D == kD, PVQA_ERR_SHAPE, "attn_fwd_v3: head dim %lld unsupported (kernel is specialised for 64)", (long long)D);
  PVQA_REQUIRE(B >= 0 && H > 0 && Sq >= 0 && Sk >= 0, PVQA_ERR_SHAPE, "attn_fwd_v3: bad dimension");
  if (B == 0 || Sq == 0) return PVQA_OK;
  PVQA_REQUIRE(Sk > 0, PVQA_ERR_SHAPE, "attn_fwd_v3: Sk must be > 0");
  PVQA_REQUIRE(q && k && v && o, PVQA_ERR_NULL, "attn_fwd_v3: NULL pointer");
  PVQA_REQUIRE(!causal || Sq == Sk, PVQA_ERR_SHAPE, "attn_fwd_v3: causal requires Sq == Sk");
  PVQA_REQUIRE(H <= 65535 && B <= 65535, PVQA_ERR_SHAPE, "attn_fwd_v3: H and B must be <= 65535");
  PVQA_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0 && o_stride_s % 8 == 0 && o_stride_h % 8 == 0 &&
                   o_stride_b % 8 == 0,
               PVQA_ERR_ALIGN, "attn_fwd_v3: output rows must be 16-byte aligned");
  const int64_t n_kpad = (Sk + kBN - 1) / kBN * kBN;
  const int64_t n_floats = n_kpad + (rel_bias ? 4 * (int64_t)pvqa_f3::rel_copy_stride((int)Sq, (int)n_kpad) + 32 : 0);
  const size_t smem_bytes = 1024 + kF3OffFloats + (size_t)n_floats * 4;
  // up to ~113 KB two CTAs share an SM; beyond that (very long sequences) the kernel still runs, one CTA per SM
  PVQA_REQUIRE(smem_bytes <= 227 * 1024, PVQA_ERR_SHAPE, "attn_fwd_v3: Sq/Sk too large for the bias staging buffers");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap(&tq, q, B, Sq, H, q_stride_b, q_stride_s, q_stride_h, kBM, "q"))) return rc;
  if ((rc = make_tmap(&tk, k, B, Sk, H, k_stride_b, k_stride_s, k_stride_h, kBN, "k"))) return rc;
  if ((rc = make_tmap(&tv, v, B, Sk, H, v_stride_b, v_stride_s, v_stride_h, kBN, "v"))) return rc;
  AttnFwdParams p{};
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.lse = lse; p.rel_bias = rel_bias; p.key_add = key_add;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.o_stride_b = o_stride_b; p.o_stride_s = o_stride_s; p.o_stride_h = o_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset; p.rng_base = g_rng_base;
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  const bool rel = rel_bias != nullptr, drop = p.drop_thr8 != 0;
  const bool scp = scp_bucket != nullptr;                 // implies rel (checked above)
  auto kern = scp   ? (drop ? attn_fwd3_kernel<true, true, true> : attn_fwd3_kernel<true, false, true>)
              : rel ? (drop ? attn_fwd3_kernel<true, true, false> : attn_fwd3_kernel<true, false, false>)
                    : (drop ? attn_fwd3_kernel<false, true, false> : attn_fwd3_kernel<false, false, false>);
  static bool attr_set[6] = {false, false, false, false, false, false};
  const int vi = (scp ? 4 : rel ? 2 : 0) + (drop ? 1 : 0);
  if (!attr_set[vi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_fwd_v3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[vi] = true;
  }
  dim3 grid((unsigned)((Sq + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  kern<<<grid, kF3Threads, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_fwd_v3");
  return PVQA_OK;
}
