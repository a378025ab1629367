// attn.cu — K2 / K3: flash attention on tcgen05 tensor cores (TMEM accumulators, TMA-fed).
//
// reference semantics restated (no code shared):
//   K2: HF T5Attention.forward core (transformers modeling_t5.py:312-336) as called from
//       core/model/PhonemeLaTr.py:111-114 — UNSCALED q.k + shared bucketed relative bias + key mask,
//       fp32 softmax, P@V.
//   K3: nn.MultiheadAttention core inside nn.TransformerDecoder (core/model/modules/transformer_utils.py:47-64,
//       core/model/PhonemeLaTr.py:134-144) — q.k/sqrt(D) + causal -inf + FLOAT (additive) key masks.
//
// One CTA = one (batch, head, 128-query tile).  Per 128-key tile:
//   TMA(K,V) -> smem (SWIZZLE_128B)          S = Q K^T   tcgen05.mma 128x128x64  -> TMEM[0,128)
//   softmax warps: tcgen05.ld S, *scale + rel_bias[j-i] + key_add[j] (+causal), online max/sum in the
//   exp2 domain, P -> bf16 -> smem (K-major SW128)            O_j = P V   tcgen05.mma 128x64x128 -> TMEM[128,192)
//   O_j is pulled to registers and merged into the running (rescaled) fp32 output.
// The T5 bias is never materialised as (H,S,S): it is a (H, Sq+Sk-1) vector over relative offsets
// held in shared memory.
#include "common.cuh"
#include "tc05.cuh"

namespace pvqa {

constexpr int kBM = 128;        // query rows per CTA (UMMA M)
constexpr int kBN = 128;        // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kD = 64;          // head dim
constexpr int kAttnThreads = 128;
constexpr uint32_t kTmemCols = 256;   // S: [0,128)  O_j: [128,192)   (power of two >= 192)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnFwdParams {
  __nv_bfloat16* o;
  float* lse;                 // (B,H,Sq)
  const float* rel_bias;      // (H, Sq+Sk-1) or null
  const float* key_add;       // (B, Sk) or null
  int B, H, Sq, Sk;
  long long o_stride_b, o_stride_s, o_stride_h;
  float scale;
  int causal;
  uint32_t drop_thr8;         // 0 = no dropout; drop probability = thr8/256
  float drop_scale;           // 1 / keep probability
  uint64_t seed, offset;
};

// smem carve-up (bytes from the 1024-aligned base)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kBM * kD * 2;            // 16 KB
constexpr int kOffV = kOffK + kBN * kD * 2;            // 32 KB
constexpr int kOffP = kOffV + kBN * kD * 2;            // 48 KB, 32 KB long (two [128][64] sub-tiles)
constexpr int kOffBar = kOffP + kBM * kBN * 2;         // 80 KB
constexpr int kOffFloats = kOffBar + 64;               // rel bias then key_add

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_s = bar_q + 2;
  uint64_t* bar_o = bar_q + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 4);
  float* s_rel = reinterpret_cast<float*>(smem + kOffFloats);
  const int n_rel = p.rel_bias ? (p.Sq + p.Sk - 1) : 0;
  float* s_kadd = s_rel + n_rel;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int i0 = blockIdx.x * kBM;
  const int h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV);
    tc05::mbar_init(bar_q, 1); tc05::mbar_init(bar_kv, 1); tc05::mbar_init(bar_s, 1); tc05::mbar_init(bar_o, 1);
    tc05::fence_barrier_init();
  }
  if (warp == 0) {
    tc05::tmem_alloc(tmem_slot, kTmemCols);
    tc05::tmem_relinquish();
  }
  // stage bias vectors (pre-multiplied by log2 e: the softmax runs in the exp2 domain)
  for (int r = tid; r < n_rel; r += kAttnThreads) s_rel[r] = p.rel_bias[(long long)h * n_rel + r] * kLog2e;
  if (p.key_add)
    for (int j = tid; j < p.Sk; j += kAttnThreads) s_kadd[j] = p.key_add[(long long)b * p.Sk + j] * kLog2e;
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes

  if (tid == 0) {
    tc05::mbar_expect_tx(bar_q, kBM * kD * 2);
    tc05::tma_load_4d(smem + kOffQ, &tmQ, bar_q, 0, h, i0, b);
  }

  const int i = i0 + tid;                         // this thread's query row
  const float sl2 = p.scale * kLog2e;
  int n_tiles = (p.Sk + kBN - 1) / kBN;
  if (p.causal) {
    const int last = min(i0 + kBM - 1, p.Sq - 1);  // largest row in this tile
    n_tiles = min(n_tiles, last / kBN + 1);
  }
  const uint32_t idesc_qk = tc05::idesc_bf16(kBM, kBN, 0, 0);
  const uint32_t idesc_pv = tc05::idesc_bf16(kBM, kD, 0, 1);      // B = V is MN-major (d contiguous)
  const uint32_t q_addr = tc05::smem_u32(smem + kOffQ), k_addr = tc05::smem_u32(smem + kOffK);
  const uint32_t v_addr = tc05::smem_u32(smem + kOffV), p_addr = tc05::smem_u32(smem + kOffP);

  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[kD];
#pragma unroll
  for (int c = 0; c < kD; ++c) o_acc[c] = 0.f;

  tc05::mbar_wait(bar_q, 0);

  for (int t = 0; t < n_tiles; ++t) {
    const int j0 = t * kBN;
    const uint32_t ph = t & 1;
    if (tid == 0) {
      tc05::mbar_expect_tx(bar_kv, 2 * kBN * kD * 2);
      tc05::tma_load_4d(smem + kOffK, &tmK, bar_kv, 0, h, j0, b);
      tc05::tma_load_4d(smem + kOffV, &tmV, bar_kv, 0, h, j0, b);
      tc05::mbar_wait(bar_kv, ph);
      tc05::tc_fence_after_sync();
      // S = Q K^T : 4 k-steps of 16 over D = 64 (32 bytes per step inside the 128-byte swizzled row)
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks) {
        const uint64_t a = tc05::smem_desc_sw128(q_addr + ks * 32, 16, 1024);
        const uint64_t bd = tc05::smem_desc_sw128(k_addr + ks * 32, 16, 1024);
        tc05::mma_bf16_ss(tmem_base, a, bd, idesc_qk, ks > 0);
      }
      tc05::mma_commit(bar_s);
    }
    tc05::mbar_wait(bar_s, ph);
    tc05::tc_fence_after_sync();

    // ---- pass 1: row max of the biased scores ----
    const int relbase = p.Sq - 1 - i;              // rel index = j + relbase
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + c * 32, r);
      tc05::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const int j = j0 + c * 32 + x;
        float s = __uint_as_float(r[x]) * sl2;
        if (n_rel) s += s_rel[min(max(j + relbase, 0), n_rel - 1)];
        if (p.key_add) s += s_kadd[min(j, p.Sk - 1)];
        if (j >= p.Sk || (p.causal && j > i)) s = -INFINITY;
        mx = fmaxf(mx, s);
      }
    }
    const float m_new = fmaxf(m_run, mx);
    const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = fast_exp2(m_run - m_safe);
    // ---- pass 2: P = exp2(s - m), row sum, bf16 P -> smem (K-major, 128B swizzle) ----
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + c * 32, r);
      tc05::tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const int j = j0 + c * 32 + x;
        float s = __uint_as_float(r[x]) * sl2;
        if (n_rel) s += s_rel[min(max(j + relbase, 0), n_rel - 1)];
        if (p.key_add) s += s_kadd[min(j, p.Sk - 1)];
        if (j >= p.Sk || (p.causal && j > i)) s = -INFINITY;
        pv[x] = fast_exp2(s - m_safe);
        sum += pv[x];
      }
      if (p.drop_thr8) {      // dropout acts on the normalised probabilities: the row sum stays undropped
        const uint64_t grow = ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * (uint64_t)((p.Sk + 15) >> 4);
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const uint32_t keep = attn_dropout_keep16(p.seed, p.offset, grow + ((j0 + c * 32) >> 4) + g2, p.drop_thr8);
#pragma unroll
          for (int x = 0; x < 16; ++x) pv[g2 * 16 + x] = ((keep >> x) & 1u) ? pv[g2 * 16 + x] * p.drop_scale : 0.f;
        }
      }
      // 32 columns = 4 chunks of 8 bf16 (16 B); sub-tile = c / 2, chunk-in-row = (c % 2) * 4 + q
      uint8_t* prow = smem + kOffP + (c >> 1) * (kBM * 128) + tid * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = f32x2_to_bf16x2(pv[q * 8 + 0], pv[q * 8 + 1]);
        u.y = f32x2_to_bf16x2(pv[q * 8 + 2], pv[q * 8 + 3]);
        u.z = f32x2_to_bf16x2(pv[q * 8 + 4], pv[q * 8 + 5]);
        u.w = f32x2_to_bf16x2(pv[q * 8 + 6], pv[q * 8 + 7]);
        const int chunk = ((c & 1) * 4 + q) ^ (tid & 7);
        *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
      }
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;
#pragma unroll
    for (int c = 0; c < kD; ++c) o_acc[c] *= alpha;

    tc05::fence_proxy_async_smem();
    tc05::tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc05::tc_fence_after_sync();
      // O_j = P V : 8 k-steps of 16 keys.  A = P (K-major: sub-tile ks/4, +32 B per step),
      // B = V (MN-major: 16 keys = 2 swizzle atoms of 8 rows x 128 B = 2048 B per step)
#pragma unroll
      for (int ks = 0; ks < kBN / 16; ++ks) {
        const uint64_t a = tc05::smem_desc_sw128(p_addr + (ks >> 2) * (kBM * 128) + (ks & 3) * 32, 16, 1024);
        const uint64_t bd = tc05::smem_desc_sw128(v_addr + ks * 2048, 16, 1024);
        tc05::mma_bf16_ss(tmem_base + kBN, a, bd, idesc_pv, ks > 0);
      }
      tc05::mma_commit(bar_o);
    }
    tc05::mbar_wait(bar_o, ph);
    tc05::tc_fence_after_sync();
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + kBN + c * 32, r);
      tc05::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) o_acc[c * 32 + x] += __uint_as_float(r[x]);
    }
    tc05::tc_fence_before_sync();
    __syncthreads();          // TMEM S/O_j and smem K/V/P are free for the next tile
    tc05::tc_fence_after_sync();
  }

  // ---- epilogue: normalise, write O (bf16) and lse (natural log) ----
  if (i < p.Sq) {
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = p.o + (long long)b * p.o_stride_b + (long long)i * p.o_stride_s + (long long)h * p.o_stride_h;
#pragma unroll
    for (int c = 0; c < kD / 8; ++c) {
      uint4 u;
      u.x = f32x2_to_bf16x2(o_acc[c * 8 + 0] * inv, o_acc[c * 8 + 1] * inv);
      u.y = f32x2_to_bf16x2(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv);
      u.z = f32x2_to_bf16x2(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv);
      u.w = f32x2_to_bf16x2(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv);
      *reinterpret_cast<uint4*>(orow + c * 8) = u;
    }
    if (p.lse)
      p.lse[((long long)b * p.H + h) * p.Sq + i] =
          l_run > 0.f ? (m_run + log2f(l_run)) * (1.0f / kLog2e) : -INFINITY;
  }
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------
// backward.  One CTA = one (batch, head, 128-key tile); loop over 128-query tiles.
//   S  = Q K^T, dP = dO V^T                       (tcgen05, TMEM [0,128) and [128,256))
//   P  = exp2(s - lse), dS = P * (dP - delta)      (256 threads: row = TMEM lane, half of the columns each)
//   dV += P^T dO, dK += dS^T Q  (A operands read MN-major from the same [query][key] smem tiles)
//   dQ_m = dS K -> TMEM (aliasing S) -> fp32 RED into the dq accumulator (other key tiles add to it)
//   d_rel[j-i] += dS : diagonal sums of the bf16 dS tile in smem, one diagonal per thread, overlapped with the MMAs
// ---------------------------------------------------------------------------------
constexpr int kBwdThreads2 = 256;
constexpr uint32_t kBwdTmemCols = 512;   // S/dQ [0,128) | dP [128,256) | dV [256,320) | dK [320,384)
constexpr int kBOffK = 0;
constexpr int kBOffV = kBOffK + kBN * kD * 2;
constexpr int kBOffQ = kBOffV + kBN * kD * 2;
constexpr int kBOffdO = kBOffQ + kBM * kD * 2;
constexpr int kBOffP = kBOffdO + kBM * kD * 2;            // 64 KB
constexpr int kBOffdS = kBOffP + kBM * kBN * 2;           // 96 KB
constexpr int kBOffBar = kBOffdS + kBM * kBN * 2;         // 128 KB
constexpr int kBOffFloats = kBOffBar + 64;

struct AttnBwdParams {
  const __nv_bfloat16* o;
  const __nv_bfloat16* d_o;
  const float* lse;
  const float* rel_bias;
  const float* key_add;
  float* dq_accum;            // (B,Sq,H,64) fp32, zero-initialised
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  float* d_rel;               // (H, Sq+Sk-1) fp32 accumulated, or null
  int B, H, Sq, Sk;
  long long o_stride_b, o_stride_s, o_stride_h;
  long long do_stride_b, do_stride_s, do_stride_h;
  long long dk_stride_b, dk_stride_s, dk_stride_h;
  long long dv_stride_b, dv_stride_s, dv_stride_h;
  float scale;
  int causal;
  uint32_t drop_thr8;
  float drop_scale;
  uint64_t seed, offset;
};

__global__ void __launch_bounds__(kBwdThreads2, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kBOffBar);
  uint64_t* bar_ld = bar_kv + 1;
  uint64_t* bar_s = bar_kv + 2;
  uint64_t* bar_dq = bar_kv + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 4);
  const int n_rel = p.rel_bias ? (p.Sq + p.Sk - 1) : 0;
  const int n_win = p.rel_bias ? (p.Sq + kBN - 1) : 0;      // rel offsets this key tile can see
  float* s_rel = reinterpret_cast<float*>(smem + kBOffFloats);   // [n_win] bias * log2e
  float* s_drel = s_rel + n_win;                                 // [n_win] gradient accumulator
  float* s_kadd = s_drel + n_win;                                // [kBN]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowl = (warp & 3) * 32 + lane;     // row inside the 128-row tile == TMEM lane
  const int half = warp >> 2;                  // which half of the columns this thread owns
  const int j0 = blockIdx.x * kBN;
  const int h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV); tc05::prefetch_tmap(&tmdO);
    tc05::mbar_init(bar_kv, 1); tc05::mbar_init(bar_ld, 1); tc05::mbar_init(bar_s, 1); tc05::mbar_init(bar_dq, 1);
    tc05::fence_barrier_init();
  }
  if (warp == 0) {
    tc05::tmem_alloc(tmem_slot, kBwdTmemCols);
    tc05::tmem_relinquish();
  }
  // window of relative offsets: global rel index r = j - i + Sq - 1 = j0 + w, w in [0, n_win)
  for (int w = tid; w < n_win; w += kBwdThreads2) {
    const int r = j0 + w;
    s_rel[w] = (r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
    s_drel[w] = 0.f;
  }
  if (tid < kBN) {
    const int j = j0 + tid;
    s_kadd[tid] = (p.key_add && j < p.Sk) ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f;
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  if (tid == 0) {
    tc05::mbar_expect_tx(bar_kv, 2 * kBN * kD * 2);
    tc05::tma_load_4d(smem + kBOffK, &tmK, bar_kv, 0, h, j0, b);
    tc05::tma_load_4d(smem + kBOffV, &tmV, bar_kv, 0, h, j0, b);
  }

  const float sl2 = p.scale * kLog2e;
  const uint32_t idesc_s = tc05::idesc_bf16(kBM, kBN, 0, 0);
  const uint32_t idesc_dkv = tc05::idesc_bf16(kBN, kD, 1, 1);   // A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
  const uint32_t idesc_dq = tc05::idesc_bf16(kBM, kD, 0, 1);    // A = dS (K-major), B = K (MN-major)
  const uint32_t k_addr = tc05::smem_u32(smem + kBOffK), v_addr = tc05::smem_u32(smem + kBOffV);
  const uint32_t q_addr = tc05::smem_u32(smem + kBOffQ), do_addr = tc05::smem_u32(smem + kBOffdO);
  const uint32_t p_addr = tc05::smem_u32(smem + kBOffP), ds_addr = tc05::smem_u32(smem + kBOffdS);

  const int m_tiles = (p.Sq + kBM - 1) / kBM;
  const int m_first = p.causal ? (j0 / kBM) : 0;      // query tiles entirely above the diagonal see nothing
  int it = 0;
  for (int mt = m_first; mt < m_tiles; ++mt, ++it) {
    const int i0 = mt * kBM;
    const uint32_t ph = it & 1;
    if (tid == 0) {
      tc05::mbar_expect_tx(bar_ld, 2 * kBM * kD * 2);
      tc05::tma_load_4d(smem + kBOffQ, &tmQ, bar_ld, 0, h, i0, b);
      tc05::tma_load_4d(smem + kBOffdO, &tmdO, bar_ld, 0, h, i0, b);
      if (it == 0) tc05::mbar_wait(bar_kv, 0);
      tc05::mbar_wait(bar_ld, ph);
      tc05::tc_fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks) {      // S = Q K^T
        tc05::mma_bf16_ss(tmem_base, tc05::smem_desc_sw128(q_addr + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(k_addr + ks * 32, 16, 1024), idesc_s, ks > 0);
      }
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks) {      // dP = dO V^T
        tc05::mma_bf16_ss(tmem_base + kBN, tc05::smem_desc_sw128(do_addr + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(v_addr + ks * 32, 16, 1024), idesc_s, ks > 0);
      }
      tc05::mma_commit(bar_s);
    }
    // ---- per-row statistics (overlaps the MMAs): lse and delta = rowsum(dO * O) ----
    const int i = i0 + rowl;
    float lse2 = 0.f, delta = 0.f;
    const bool row_ok = i < p.Sq;
    if (row_ok) {
      lse2 = p.lse[((long long)b * p.H + h) * p.Sq + i] * kLog2e;
      const __nv_bfloat16* orow = p.o + (long long)b * p.o_stride_b + (long long)i * p.o_stride_s + (long long)h * p.o_stride_h;
      const __nv_bfloat16* grow = p.d_o + (long long)b * p.do_stride_b + (long long)i * p.do_stride_s + (long long)h * p.do_stride_h;
#pragma unroll
      for (int c = 0; c < kD / 8; ++c) {
        f8 a = Vec8<__nv_bfloat16>::load(orow + c * 8);
        f8 g = Vec8<__nv_bfloat16>::load(grow + c * 8);
#pragma unroll
        for (int x = 0; x < 8; ++x) delta = fmaf(a.v[x], g.v[x], delta);
      }
    }
    const bool row_live = row_ok && (lse2 != -INFINITY);
    tc05::mbar_wait(bar_s, ph);
    tc05::tc_fence_after_sync();

    // ---- P and dS for this thread's 64 columns (sub-tile `half`) ----
    const int wbase = p.Sq - 1 - i;                  // window index = jl + wbase  (jl = j - j0)
    uint8_t* prow = smem + kBOffP + half * (kBM * 128) + rowl * 128;
    uint8_t* dsrow = smem + kBOffdS + half * (kBM * 128) + rowl * 128;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t rs[32], rp[32];
      tc05::tmem_ld_32x32(tmem_row + half * 64 + c * 32, rs);
      tc05::tmem_ld_32x32(tmem_row + kBN + half * 64 + c * 32, rp);
      tc05::tmem_ld_wait();
      float pv[32], dsv[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const int jl = half * 64 + c * 32 + x;
        const int j = j0 + jl;
        float s = __uint_as_float(rs[x]) * sl2 + s_kadd[jl];
        if (n_win) s += s_rel[min(max(jl + wbase, 0), n_win - 1)];
        const bool dead = !row_live || j >= p.Sk || (p.causal && j > i);
        const float pr = dead ? 0.f : fast_exp2(s - lse2);
        pv[x] = pr;
        dsv[x] = __uint_as_float(rp[x]);
      }
      if (p.drop_thr8) {      // P_drop = P*M/keep feeds dV; dP = M/keep * dP_drop feeds dS
        const uint64_t grow = ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * (uint64_t)((p.Sk + 15) >> 4);
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const uint32_t keep = attn_dropout_keep16(p.seed, p.offset,
                                                    grow + ((j0 + half * 64 + c * 32) >> 4) + g2, p.drop_thr8);
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float mk = ((keep >> x) & 1u) ? p.drop_scale : 0.f;
            const float pr = pv[g2 * 16 + x];
            dsv[g2 * 16 + x] = pr * (mk * dsv[g2 * 16 + x] - delta) * p.scale;
            pv[g2 * 16 + x] = pr * mk;
          }
        }
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x) dsv[x] = pv[x] * (dsv[x] - delta) * p.scale;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u, w;
        u.x = f32x2_to_bf16x2(pv[q * 8 + 0], pv[q * 8 + 1]);  u.y = f32x2_to_bf16x2(pv[q * 8 + 2], pv[q * 8 + 3]);
        u.z = f32x2_to_bf16x2(pv[q * 8 + 4], pv[q * 8 + 5]);  u.w = f32x2_to_bf16x2(pv[q * 8 + 6], pv[q * 8 + 7]);
        w.x = f32x2_to_bf16x2(dsv[q * 8 + 0], dsv[q * 8 + 1]); w.y = f32x2_to_bf16x2(dsv[q * 8 + 2], dsv[q * 8 + 3]);
        w.z = f32x2_to_bf16x2(dsv[q * 8 + 4], dsv[q * 8 + 5]); w.w = f32x2_to_bf16x2(dsv[q * 8 + 6], dsv[q * 8 + 7]);
        const int chunk = (c * 4 + q) ^ (rowl & 7);
        *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
        *reinterpret_cast<uint4*>(dsrow + chunk * 16) = w;
      }
    }
    tc05::fence_proxy_async_smem();
    tc05::tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc05::tc_fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < kBM / 16; ++ks) {     // dV += P^T dO ; dK += dS^T Q   (K = 128 query rows, 16 per step)
        const uint64_t ap = tc05::smem_desc_sw128(p_addr + ks * 2048, kBM * 128, 1024);
        const uint64_t bdo = tc05::smem_desc_sw128(do_addr + ks * 2048, 16, 1024);
        tc05::mma_bf16_ss(tmem_base + 256, ap, bdo, idesc_dkv, (it > 0) || (ks > 0));
      }
#pragma unroll
      for (int ks = 0; ks < kBM / 16; ++ks) {
        const uint64_t ads = tc05::smem_desc_sw128(ds_addr + ks * 2048, kBM * 128, 1024);
        const uint64_t bq = tc05::smem_desc_sw128(q_addr + ks * 2048, 16, 1024);
        tc05::mma_bf16_ss(tmem_base + 320, ads, bq, idesc_dkv, (it > 0) || (ks > 0));
      }
#pragma unroll
      for (int ks = 0; ks < kBN / 16; ++ks) {     // dQ_m = dS K    (K = 128 keys)
        const uint64_t ads = tc05::smem_desc_sw128(ds_addr + (ks >> 2) * (kBM * 128) + (ks & 3) * 32, 16, 1024);
        const uint64_t bk = tc05::smem_desc_sw128(k_addr + ks * 2048, 16, 1024);
        tc05::mma_bf16_ss(tmem_base, ads, bk, idesc_dq, ks > 0);
      }
      tc05::mma_commit(bar_dq);
    }
    // ---- d_rel: one diagonal of the dS tile per thread (generic-proxy reads, overlapping the MMAs) ----
    if (p.d_rel && tid < 2 * kBM - 1) {
      const int delta_ij = tid - (kBM - 1);          // jl - il
      float acc = 0.f;
      const int il_lo = max(0, -delta_ij), il_hi = min(kBM - 1, kBN - 1 - delta_ij);
      for (int il = il_lo; il <= il_hi; ++il) {
        const int jl = il + delta_ij;
        const uint8_t* e = smem + kBOffdS + (jl >> 6) * (kBM * 128) + il * 128 +
                           ((((jl & 63) >> 3) ^ (il & 7)) << 4) + (jl & 7) * 2;
        acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(e));
      }
      const int w = delta_ij - i0 + p.Sq - 1;        // window index of rel = (j0+jl) - (i0+il) + Sq-1, minus j0
      if (w >= 0 && w < n_win) s_drel[w] += acc;     // each w is owned by exactly one thread per step
    }
    tc05::mbar_wait(bar_dq, ph);
    tc05::tc_fence_after_sync();
    {
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + half * 32, r);
      tc05::tmem_ld_wait();
      if (row_ok) {
        float* dst = p.dq_accum + (((long long)b * p.Sq + i) * p.H + h) * kD + half * 32;
#pragma unroll
        for (int x = 0; x < 32; x += 4)
          red_add_v4(dst + x, __uint_as_float(r[x]), __uint_as_float(r[x + 1]), __uint_as_float(r[x + 2]),
                     __uint_as_float(r[x + 3]));
      }
    }
    tc05::tc_fence_before_sync();
    __syncthreads();
    tc05::tc_fence_after_sync();
  }

  // ---- epilogue: dV, dK rows (key j0 + rowl), columns [32*half, +32) ----
  if (it > 0) {
    const int j = j0 + rowl;
    uint32_t rv[32], rk[32];
    tc05::tmem_ld_32x32(tmem_row + 256 + half * 32, rv);
    tc05::tmem_ld_32x32(tmem_row + 320 + half * 32, rk);
    tc05::tmem_ld_wait();
    if (j < p.Sk) {
      __nv_bfloat16* dvrow = p.dv + (long long)b * p.dv_stride_b + (long long)j * p.dv_stride_s + (long long)h * p.dv_stride_h + half * 32;
      __nv_bfloat16* dkrow = p.dk + (long long)b * p.dk_stride_b + (long long)j * p.dk_stride_s + (long long)h * p.dk_stride_h + half * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u, w;
        u.x = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 0]), __uint_as_float(rv[c * 8 + 1]));
        u.y = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 2]), __uint_as_float(rv[c * 8 + 3]));
        u.z = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 4]), __uint_as_float(rv[c * 8 + 5]));
        u.w = f32x2_to_bf16x2(__uint_as_float(rv[c * 8 + 6]), __uint_as_float(rv[c * 8 + 7]));
        w.x = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 0]), __uint_as_float(rk[c * 8 + 1]));
        w.y = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 2]), __uint_as_float(rk[c * 8 + 3]));
        w.z = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 4]), __uint_as_float(rk[c * 8 + 5]));
        w.w = f32x2_to_bf16x2(__uint_as_float(rk[c * 8 + 6]), __uint_as_float(rk[c * 8 + 7]));
        *reinterpret_cast<uint4*>(dvrow + c * 8) = u;
        *reinterpret_cast<uint4*>(dkrow + c * 8) = w;
      }
    }
  } else {
    // causal tile with no visible query rows cannot happen (the diagonal tile always exists); keep outputs defined
    const int j = j0 + rowl;
    if (j < p.Sk) {
      __nv_bfloat16* dvrow = p.dv + (long long)b * p.dv_stride_b + (long long)j * p.dv_stride_s + (long long)h * p.dv_stride_h + half * 32;
      __nv_bfloat16* dkrow = p.dk + (long long)b * p.dk_stride_b + (long long)j * p.dk_stride_s + (long long)h * p.dk_stride_h + half * 32;
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4*>(dvrow + c * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dkrow + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  if (p.d_rel) {
    const float inv_scale = 1.0f / p.scale;          // the smem tile holds scale * dS
    for (int w = tid; w < n_win; w += kBwdThreads2) {
      const int r = j0 + w;
      const float g = s_drel[w];
      if (r < n_rel && g != 0.f) atomicAdd(p.d_rel + (long long)h * n_rel + r, g * inv_scale);
    }
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem_base, kBwdTmemCols);
}

// ---------------------------------------------------------------------------------
// host side: TMA descriptors
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// (B, S, H, D=64) bf16 view with element strides -> 4-D map (D, H, S, B), box {64, 1, rows, 1}, 128B swizzle
static int make_tmap(CUtensorMap* m, const void* ptr, int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss,
                     int64_t sh, int rows, const char* what) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PVQA_ERR_CUDA, "attn: cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (sb * 2) % 16 || (ss * 2) % 16 || (sh * 2) % 16)
    return fail(PVQA_ERR_ALIGN, "attn: %s base/strides must be 16-byte aligned", what);
  cuuint64_t gdim[4] = {(cuuint64_t)kD, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)sh * 2, (cuuint64_t)ss * 2, (cuuint64_t)(B > 1 ? sb : ss * S) * 2};
  cuuint32_t box[4] = {(cuuint32_t)kD, 1, (cuuint32_t)rows, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PVQA_ERR_CUDA, "attn: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return PVQA_OK;
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                             const float* rel_bias, const float* key_add, int64_t B, int64_t H, int64_t Sq,
                             int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                             int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t v_stride_b,
                             int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b, int64_t o_stride_s,
                             int64_t o_stride_h, float scale, int causal, float dropout_p, uint64_t seed,
                             uint64_t offset, void* stream) {
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "attn_fwd: dropout_p must be in [0,1)");
  PVQA_REQUIRE(D == kD, PVQA_ERR_SHAPE, "attn_fwd: head dim %lld unsupported (kernel is specialised for 64)", (long long)D);
  PVQA_REQUIRE(B >= 0 && H > 0 && Sq >= 0 && Sk >= 0, PVQA_ERR_SHAPE, "attn_fwd: bad dimension");
  if (B == 0 || Sq == 0) return PVQA_OK;
  PVQA_REQUIRE(Sk > 0, PVQA_ERR_SHAPE, "attn_fwd: Sk must be > 0");
  PVQA_REQUIRE(q && k && v && o, PVQA_ERR_NULL, "attn_fwd: NULL pointer");
  PVQA_REQUIRE(!causal || Sq == Sk, PVQA_ERR_SHAPE, "attn_fwd: causal requires Sq == Sk");
  PVQA_REQUIRE(H <= 65535 && B <= 65535, PVQA_ERR_SHAPE, "attn_fwd: H and B must be <= 65535");
  PVQA_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0 && o_stride_s % 8 == 0 && o_stride_h % 8 == 0 &&
                   o_stride_b % 8 == 0,
               PVQA_ERR_ALIGN, "attn_fwd: output rows must be 16-byte aligned");
  const int64_t n_floats = (rel_bias ? Sq + Sk - 1 : 0) + (key_add ? Sk : 0);
  const size_t smem_bytes = 1024 + kOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 200 * 1024, PVQA_ERR_SHAPE, "attn_fwd: Sq/Sk too large for the bias staging buffer");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap(&tq, q, B, Sq, H, q_stride_b, q_stride_s, q_stride_h, kBM, "q"))) return rc;
  if ((rc = make_tmap(&tk, k, B, Sk, H, k_stride_b, k_stride_s, k_stride_h, kBN, "k"))) return rc;
  if ((rc = make_tmap(&tv, v, B, Sk, H, v_stride_b, v_stride_s, v_stride_h, kBN, "v"))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  AttnFwdParams p{};
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.lse = lse; p.rel_bias = rel_bias; p.key_add = key_add;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.o_stride_b = o_stride_b; p.o_stride_s = o_stride_s; p.o_stride_h = o_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset;
  dim3 grid((unsigned)((Sq + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<<<grid, kAttnThreads, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_fwd");
  return PVQA_OK;
}


extern "C" int pvqa_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                             const float* lse, const float* rel_bias, const float* key_add, float* dq_accum,
                             void* dk, void* dv, float* d_rel_bias, int64_t B, int64_t H, int64_t Sq, int64_t Sk,
                             int64_t D, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                             int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t v_stride_b,
                             int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b, int64_t o_stride_s,
                             int64_t o_stride_h, int64_t do_stride_b, int64_t do_stride_s, int64_t do_stride_h,
                             int64_t dk_stride_b, int64_t dk_stride_s, int64_t dk_stride_h, int64_t dv_stride_b,
                             int64_t dv_stride_s, int64_t dv_stride_h, float scale, int causal, float dropout_p,
                             uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "attn_bwd: dropout_p must be in [0,1)");
  PVQA_REQUIRE(D == kD, PVQA_ERR_SHAPE, "attn_bwd: head dim %lld unsupported (kernel is specialised for 64)", (long long)D);
  PVQA_REQUIRE(B >= 0 && H > 0 && Sq >= 0 && Sk >= 0, PVQA_ERR_SHAPE, "attn_bwd: bad dimension");
  if (B == 0 || Sq == 0 || Sk == 0) return PVQA_OK;
  PVQA_REQUIRE(q && k && v && o && d_o && lse && dq_accum && dk && dv, PVQA_ERR_NULL, "attn_bwd: NULL pointer");
  PVQA_REQUIRE(!causal || Sq == Sk, PVQA_ERR_SHAPE, "attn_bwd: causal requires Sq == Sk");
  PVQA_REQUIRE(scale != 0.f, PVQA_ERR_SHAPE, "attn_bwd: scale must be non-zero");
  PVQA_REQUIRE(!d_rel_bias || rel_bias, PVQA_ERR_NULL, "attn_bwd: d_rel_bias requested without rel_bias");
  PVQA_REQUIRE(H <= 65535 && B <= 65535, PVQA_ERR_SHAPE, "attn_bwd: H and B must be <= 65535");
  auto al8 = [](int64_t a, int64_t b2, int64_t c) { return a % 8 == 0 && b2 % 8 == 0 && c % 8 == 0; };
  PVQA_REQUIRE(aligned16(o) && aligned16(d_o) && aligned16(dk) && aligned16(dv) && aligned16(dq_accum) &&
                   al8(o_stride_b, o_stride_s, o_stride_h) && al8(do_stride_b, do_stride_s, do_stride_h) &&
                   al8(dk_stride_b, dk_stride_s, dk_stride_h) && al8(dv_stride_b, dv_stride_s, dv_stride_h),
               PVQA_ERR_ALIGN, "attn_bwd: rows must be 16-byte aligned");
  const int64_t n_floats = (rel_bias ? 2 * (Sq + kBN - 1) : 0) + kBN;
  const size_t smem_bytes = 1024 + kBOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 220 * 1024, PVQA_ERR_SHAPE, "attn_bwd: Sq too large for the bias window buffers");
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap(&tq, q, B, Sq, H, q_stride_b, q_stride_s, q_stride_h, kBM, "q"))) return rc;
  if ((rc = make_tmap(&tk, k, B, Sk, H, k_stride_b, k_stride_s, k_stride_h, kBN, "k"))) return rc;
  if ((rc = make_tmap(&tv, v, B, Sk, H, v_stride_b, v_stride_s, v_stride_h, kBN, "v"))) return rc;
  if ((rc = make_tmap(&tdo, d_o, B, Sq, H, do_stride_b, do_stride_s, do_stride_h, kBM, "d_o"))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  AttnBwdParams p{};
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
  p.lse = lse; p.rel_bias = rel_bias; p.key_add = key_add; p.dq_accum = dq_accum;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.d_rel = d_rel_bias;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.o_stride_b = o_stride_b; p.o_stride_s = o_stride_s; p.o_stride_h = o_stride_h;
  p.do_stride_b = do_stride_b; p.do_stride_s = do_stride_s; p.do_stride_h = do_stride_h;
  p.dk_stride_b = dk_stride_b; p.dk_stride_s = dk_stride_s; p.dk_stride_h = dk_stride_h;
  p.dv_stride_b = dv_stride_b; p.dv_stride_s = dv_stride_s; p.dv_stride_h = dv_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset;
  dim3 grid((unsigned)((Sk + kBN - 1) / kBN), (unsigned)H, (unsigned)B);
  attn_bwd_kernel<<<grid, kBwdThreads2, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, tdo, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_bwd");
  return PVQA_OK;
}
