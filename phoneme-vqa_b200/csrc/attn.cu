// attn.cu — K2 / K3: flash attention on tcgen05 tensor cores (TMEM accumulators, TMA-fed).
//
// reference semantics restated (no code shared):
//   K2: HF T5Attention.forward core (transformers modeling_t5.py:312-336) as called from
//       core/model/PhonemeLaTr.py:111-114 — UNSCALED q.k + shared bucketed relative bias + key mask,
//       fp32 softmax, dropout on P, P@V.
//   K3: nn.MultiheadAttention core inside nn.TransformerDecoder (core/model/modules/transformer_utils.py:47-64,
//       core/model/PhonemeLaTr.py:134-144) — q.k/sqrt(D) + causal -inf + FLOAT (additive) key masks.
//
// Forward: one CTA = one (batch, head, 128-query tile), 256 threads, 2 CTAs per SM.  Per 128-key tile
//   TMA(K,V) -> smem (SWIZZLE_128B);  S = Q K^T  (tcgen05.mma 128x128x64 -> TMEM[0,128))
//   pass A: tcgen05.ld S, s = acc*scale + rel_bias[j-i] + key_add[j] (exp2 domain), row max, s written back to TMEM
//   pass B: p = exp2(s - m) (dropout folded in), bf16 P -> smem (K-major SW128);  O_j = P V (tcgen05.mma -> TMEM[128,192))
//   O_j merged into the rescaled fp32 running output held in registers.
// Every row is shared by two threads (TMEM lane = row, each thread owns half of the columns); they exchange the
// row max through shared memory.  The T5 bias is never materialised as (H,S,S): it is a (H, Sq+Sk-1) vector over
// relative offsets staged in shared memory; the key tail (j >= Sk) is a -inf entry of the staged key_add vector,
// so the inner loops carry no bounds checks.
// Roofline note: at D = 64 with bias + dropout the softmax costs ~13-17 ALU/SFU instructions per
// score against 2*2*64 tensor flops, so these kernels are issue-bound on the CUDA cores long before the
// tensor pipe saturates; the reported tensor-pipe fraction is what that ceiling allows.
#include "common.cuh"
#include "tc05.cuh"

namespace pvqa {

constexpr int kBM = 128;        // query rows per tile (UMMA M)
constexpr int kBN = 128;        // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kD = 64;          // head dim
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = f32x2_to_bf16x2(v[0], v[1]); u.y = f32x2_to_bf16x2(v[2], v[3]);
  u.z = f32x2_to_bf16x2(v[4], v[5]); u.w = f32x2_to_bf16x2(v[6], v[7]);
  return u;
}
// dropout keep-masks for 16 consecutive keys as four words of per-byte 0xFF/0x00 (SIMD byte compare)
__device__ __forceinline__ uint4 keep_bytes16(uint64_t seed, uint64_t offset, uint64_t group, uint32_t thr4) {
  const uint64_t c = offset + group;
  uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5A17u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return make_uint4(__vcmpgeu4(r.x, thr4), __vcmpgeu4(r.y, thr4), __vcmpgeu4(r.z, thr4), __vcmpgeu4(r.w, thr4));
}
// byte `bb` of m replicated to a full-word mask
#define PVQA_BYTE_MASK(m, bb) __byte_perm((m), 0u, 0x1111u * (bb))

// Developer-only phase trace (tools/attn_trace.py builds a private copy of the library with -DPVQA_ATTN_TRACE):
// CTAs (x=1, y=3, z<64) record clock64() at phase boundaries, thread 0 in slots [0,32), the first lane of the
// last warp in [32,64).  Compiled out of the product library.
#ifdef PVQA_ATTN_TRACE
__device__ long long g_attn_trace[64 * 64];
#define PVQA_TRACE(ev)                                                                              \
  do {                                                                                              \
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z < 64 && (ev) < 32) {                       \
      if (threadIdx.x == 0) g_attn_trace[blockIdx.z * 64 + (ev)] = clock64();                       \
      if (threadIdx.x == blockDim.x - 32) g_attn_trace[blockIdx.z * 64 + 32 + (ev)] = clock64();    \
    }                                                                                               \
  } while (0)
#else
#define PVQA_TRACE(ev)
#endif

// =================================================================================
// forward
// =================================================================================
constexpr int kFwdThreads = 256;
constexpr uint32_t kTmemCols = 256;   // S: [0,128)  O_j: [128,192)   (power of two >= 192)
constexpr int kOffQ = 0;
constexpr int kOffKV = kOffQ + kBM * kD * 2;           // 16 KB: three rotating 16 KB buffers for K_j, V_j, K_{j+1}
constexpr int kKVBuf = kBN * kD * 2;
constexpr int kOffP = kOffKV + 3 * kKVBuf;             // 64 KB, 32 KB long (two [128][64] sub-tiles)
constexpr int kOffBar = kOffP + kBM * kBN * 2;         // 96 KB
constexpr int kOffXchg = kOffBar + 64;                 // [2][128] floats
constexpr int kOffFloats = kOffXchg + 2 * kBM * 4;     // key_add (padded) then rel bias (padded)
constexpr int kRelPad = 128;

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
  const unsigned long long* rng_base;
  // SaL spatial (SCP) bias: bias += scp_tab[h][scp_bucket[b][i-q0][j-q0]] on the OCR x OCR block
  const uint8_t* scp_bucket;  // (B, L, L) or null
  const float* scp_tab;       // (H, 32)
  int scp_q0, scp_L;
};

// 32 bucket ids (bytes) of one row chunk -> add the staged per-head table values; chunk starts are multiples of 16
// relative to the block (host checks q0 % 16 == 0 and L % 16 == 0), so each 16-byte half is inside or outside.
__device__ __forceinline__ void load_scp32(const uint8_t* row, int jj0, int L, const float* s_tab, float (&out)[32]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int jj = jj0 + hf * 16;
    if (jj >= 0 && jj < L) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + jj));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) out[hf * 16 + k] = s_tab[(w[k >> 2] >> (8 * (k & 3))) & 31u];
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) out[hf * 16 + k] = 0.f;
    }
  }
}

template <bool HAS_REL, bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an OFFSET into the __shared__ array so that the compiler
  // keeps the shared address space (32-bit LDS/STS instead of generic 64-bit LD/ST for every smem access)
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_k = bar_q + 1;            // [2] by tile parity
  uint64_t* bar_v = bar_q + 3;            // [2]
  uint64_t* bar_s = bar_q + 5;
  uint64_t* bar_o = bar_q + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 7);
  float* s_x = reinterpret_cast<float*>(smem + kOffXchg);
  const int n_tiles_all = (p.Sk + kBN - 1) / kBN;
  const int n_kpad = n_tiles_all * kBN;
  float* s_kadd = reinterpret_cast<float*>(smem + kOffFloats);       // [n_kpad], -inf beyond Sk
  float* s_rel = s_kadd + n_kpad;                                     // [kRelPad + Sq + n_kpad], index r + kRelPad
  float* s_scp = s_rel + (HAS_REL ? kRelPad + p.Sq + n_kpad : 0);     // [32] SCP table of this head (SaL)
  const int n_rel = p.Sq + p.Sk - 1;
  const bool has_scp = HAS_REL && p.scp_bucket != nullptr;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowl = (warp & 3) * 32 + lane;          // row in the tile == TMEM lane
  const int half = warp >> 2;                       // column half owned by this thread
  PVQA_TRACE(0);
  const int i0 = blockIdx.x * kBM;
  const int h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV);
    tc05::mbar_init(bar_q, 1); tc05::mbar_init(bar_k, 1); tc05::mbar_init(bar_k + 1, 1);
    tc05::mbar_init(bar_v, 1); tc05::mbar_init(bar_v + 1, 1); tc05::mbar_init(bar_s, 1); tc05::mbar_init(bar_o, 1);
    tc05::fence_barrier_init();
    // the first loads go out before anything else so that they overlap the bias staging below
    // (K_j lives in buffer (2j) % 3, V_j in (2j+1) % 3: K_{j+1} reuses V_{j-1}'s buffer, V_{j+1} reuses K_j's)
    tc05::mbar_expect_tx(bar_q, kBM * kD * 2);
    tc05::tma_load_4d(smem + kOffQ, &tmQ, bar_q, 0, h, i0, b);
    tc05::mbar_expect_tx(bar_k, kKVBuf);
    tc05::tma_load_4d(smem + kOffKV, &tmK, bar_k, 0, h, 0, b);
    tc05::mbar_expect_tx(bar_v, kKVBuf);
    tc05::tma_load_4d(smem + kOffKV + kKVBuf, &tmV, bar_v, 0, h, 0, b);
  }
  __syncwarp();
  if (warp == 0) {
    tc05::tmem_alloc(tmem_slot, kTmemCols);
    tc05::tmem_relinquish();
  }
  // stage the additive vectors, pre-multiplied by log2(e): the softmax runs in the exp2 domain.
  // Four independent global loads per thread and trip (the loops are latency-, not bandwidth-bound).
  for (int j0s = tid; j0s < n_kpad; j0s += 4 * kFwdThreads) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0s + u * kFwdThreads;
      v[u] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0s + u * kFwdThreads < n_kpad) s_kadd[j0s + u * kFwdThreads] = v[u];
  }
  if (HAS_REL) {
    const int n = kRelPad + p.Sq + n_kpad;
    for (int x0 = tid; x0 < n; x0 += 4 * kFwdThreads) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = x0 + u * kFwdThreads - kRelPad;
        v[u] = (r >= 0 && r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (x0 + u * kFwdThreads < n) s_rel[x0 + u * kFwdThreads] = v[u];
    }
    if (has_scp && tid < 32) s_scp[tid] = p.scp_tab[h * 32 + tid] * kLog2e;
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  PVQA_TRACE(1);

  int n_tiles = n_tiles_all;
  if (p.causal) n_tiles = min(n_tiles, min(i0 + kBM - 1, p.Sq - 1) / kBN + 1);

  const int i = i0 + rowl;
  const bool rows_dead = i0 + (warp & 3) * 32 >= p.Sq;      // all 32 query rows of this warp are past the end
  const float sl2 = p.scale * kLog2e;
  const uint32_t idesc_qk = tc05::idesc_bf16(kBM, kBN, 0, 0);
  const uint32_t idesc_pv = tc05::idesc_bf16(kBM, kD, 0, 1);      // B = V is MN-major (d contiguous)
  const uint32_t q_addr = tc05::smem_u32(smem + kOffQ), kv_addr = tc05::smem_u32(smem + kOffKV);
  const uint32_t p_addr = tc05::smem_u32(smem + kOffP);
  const float* relrow = s_rel + (p.Sq - 1 - i) + kRelPad;         // relrow[j] = bias of key j for this row
  // with dropout the 1/keep factor is folded into the exponent: p' = p/keep, row sums carry the same factor
  const float m_shift = DROP ? log2f(p.drop_scale) : 0.f;
  const uint32_t thr4 = p.drop_thr8 * 0x01010101u;
  const uint64_t drop_row = ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * (uint64_t)((p.Sk + 15) >> 4);
  const uint64_t rng_off = p.offset + ((DROP && p.rng_base) ? *p.rng_base : 0ull);
  const uint8_t* scp_row = nullptr;                 // this row's bucket ids inside the OCR block, if it is in it
  if (has_scp && i >= p.scp_q0 && i < p.scp_q0 + p.scp_L && i < p.Sq)
    scp_row = p.scp_bucket + ((long long)b * p.scp_L + (i - p.scp_q0)) * p.scp_L;

  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) o_acc[c] = 0.f;

  if (tid == 0) tc05::mbar_wait(bar_q, 0);

  for (int t = 0; t < n_tiles; ++t) {
    const int j0 = t * kBN;
    const uint32_t ph = t & 1;
    const uint32_t k_addr = kv_addr + ((2 * t) % 3) * kKVBuf;
    const uint32_t v_addr = kv_addr + ((2 * t + 1) % 3) * kKVBuf;
    if (tid == 0) {
      if (t + 1 < n_tiles) {        // prefetch K_{j+1} into the buffer V_{j-1} vacated at the end of the last tile
        tc05::mbar_expect_tx(bar_k + ((t + 1) & 1), kKVBuf);
        tc05::tma_load_4d(smem + kOffKV + ((2 * t + 2) % 3) * kKVBuf, &tmK, bar_k + ((t + 1) & 1), 0, h, j0 + kBN, b);
      }
      tc05::mbar_wait(bar_k + (t & 1), (t >> 1) & 1);
      tc05::tc_fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks)      // S = Q K^T : 32 bytes per k-step inside the 128-byte swizzled row
        tc05::mma_bf16_ss(tmem_base, tc05::smem_desc_sw128(q_addr + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(k_addr + ks * 32, 16, 1024), idesc_qk, ks > 0);
      tc05::mma_commit(bar_s);
    }
    const bool diag = p.causal && (j0 + kBN - 1 > i0);     // tile touches the diagonal (CTA-uniform)
    tc05::mbar_wait(bar_s, ph);
    tc05::tc_fence_after_sync();
    PVQA_TRACE(2 + 6 * t);
    if (tid == 0 && t + 1 < n_tiles) {   // K_j is consumed: its buffer takes V_{j+1}
      tc05::mbar_expect_tx(bar_v + ((t + 1) & 1), kKVBuf);
      tc05::tma_load_4d(smem + kOffKV + ((2 * t) % 3) * kKVBuf, &tmV, bar_v + ((t + 1) & 1), 0, h, j0 + kBN, b);
    }

    // ---- pass A: biased scores (written back to TMEM) and the row max over this thread's 64 columns ----
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int jb = j0 + half * 64 + c * 32;
      if (jb >= p.Sk || rows_dead) continue;        // warp-uniform: chunk beyond the last key / no live query row
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + half * 64 + c * 32, r);
      tc05::tmem_ld_wait();
      if (has_scp && scp_row != nullptr && jb + 32 > p.scp_q0 && jb < p.scp_q0 + p.scp_L) {
        float sb[32];
        load_scp32(scp_row, jb - p.scp_q0, p.scp_L, s_scp, sb);
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const float s = fmaf(__uint_as_float(r[x]), sl2, s_kadd[jb + x] + relrow[jb + x] + sb[x]);
          mx = fmaxf(mx, s);
          r[x] = __float_as_uint(s);
        }
      } else if (diag) {
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          float bias = s_kadd[jb + x];
          if (HAS_REL) bias += relrow[jb + x];
          float s = fmaf(__uint_as_float(r[x]), sl2, bias);
          if (jb + x > i) s = -INFINITY;
          mx = fmaxf(mx, s);
          r[x] = __float_as_uint(s);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          float bias = s_kadd[jb + x];
          if (HAS_REL) bias += relrow[jb + x];
          const float s = fmaf(__uint_as_float(r[x]), sl2, bias);
          mx = fmaxf(mx, s);
          r[x] = __float_as_uint(s);
        }
      }
      tc05::tmem_st_32x32(tmem_row + half * 64 + c * 32, r);
    }
    tc05::tmem_st_wait();
    s_x[half * kBM + rowl] = mx;
    PVQA_TRACE(3 + 6 * t);
    __syncthreads();
    PVQA_TRACE(4 + 6 * t);
    mx = fmaxf(mx, s_x[(half ^ 1) * kBM + rowl]);
    const float m_new = fmaxf(m_run, mx);
    const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = fast_exp2(m_run - m_safe);
    const float m_sub = m_safe - m_shift;
    // ---- pass B: p = exp2(s - m), partial row sum, bf16 P -> smem (K-major, 128B swizzle), sub-tile = half ----
    float sum = 0.f;
    uint8_t* prow = smem + kOffP + half * (kBM * 128) + rowl * 128;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      if (j0 + half * 64 + c * 32 >= p.Sk || rows_dead) {      // dead chunk: P must still be zero for the PV MMA
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(prow + (((c * 4 + q) ^ (rowl & 7)) * 16)) = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + half * 64 + c * 32, r);
      tc05::tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        pv[x] = fast_exp2(__uint_as_float(r[x]) - m_sub);
        sum += pv[x];
      }
      if (DROP) {
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const uint4 kb = keep_bytes16(p.seed, rng_off, drop_row + ((j0 + half * 64 + c * 32) >> 4) + g2, thr4);
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
    PVQA_TRACE(5 + 6 * t);

    tc05::fence_proxy_async_smem();
    tc05::tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc05::mbar_wait(bar_v + (t & 1), (t >> 1) & 1);
      tc05::tc_fence_after_sync();
      // O_j = P V : 8 k-steps of 16 keys.  A = P (K-major: sub-tile ks/4, +32 B per step),
      // B = V (MN-major: 16 keys = 2 swizzle atoms of 8 rows x 128 B = 2048 B per step)
#pragma unroll
      for (int ks = 0; ks < kBN / 16; ++ks)
        tc05::mma_bf16_ss(tmem_base + kBN,
                          tc05::smem_desc_sw128(p_addr + (ks >> 2) * (kBM * 128) + (ks & 3) * 32, 16, 1024),
                          tc05::smem_desc_sw128(v_addr + ks * 2048, 16, 1024), idesc_pv, ks > 0);
      tc05::mma_commit(bar_o);
    }
    tc05::mbar_wait(bar_o, ph);
    tc05::tc_fence_after_sync();
    PVQA_TRACE(6 + 6 * t);
    {
      uint32_t r[32];
      tc05::tmem_ld_32x32(tmem_row + kBN + half * 32, r);
      tc05::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) o_acc[x] = fmaf(o_acc[x], alpha, __uint_as_float(r[x]));
    }
    tc05::tc_fence_before_sync();
    __syncthreads();          // TMEM S/O_j and smem K/V/P are free for the next tile
    tc05::tc_fence_after_sync();
    PVQA_TRACE(7 + 6 * t);
  }

  // ---- epilogue: combine the two half-row sums, normalise, write O (bf16) and lse (natural log) ----
  s_x[half * kBM + rowl] = l_run;
  __syncthreads();
  const float l_tot = l_run + s_x[(half ^ 1) * kBM + rowl];      // carries the 1/keep factor when DROP
  if (i < p.Sq) {
    const float inv = l_tot > 0.f ? (DROP ? p.drop_scale : 1.f) / l_tot : 0.f;
    __nv_bfloat16* orow = p.o + (long long)b * p.o_stride_b + (long long)i * p.o_stride_s +
                          (long long)h * p.o_stride_h + half * 32;
#pragma unroll
    for (int x = 0; x < 32; ++x) o_acc[x] *= inv;
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = pack8(o_acc + c * 8);
    if (p.lse && half == 0)
      p.lse[((long long)b * p.H + h) * p.Sq + i] =
          l_tot > 0.f ? (m_run + log2f(l_tot) - m_shift) * (1.0f / kLog2e) : -INFINITY;
  }
  PVQA_TRACE(30);
  tc05::tc_fence_before_sync();
  __syncthreads();
  PVQA_TRACE(31);
  if (warp == 0) tc05::tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================
// backward
//   prep kernel: delta[b,h,i] = rowsum(dO * O)
//   main kernel: one CTA = one (batch, head, 128-key tile), 512 threads (row = TMEM lane, a quarter of the
//   columns per thread); loop over 128-query tiles with a 3-stage TMA ring of Q / dO:
//     S = Q K^T, dP = dO V^T                  (TMEM [0,128) and [128,256))
//     P = exp2(s - lse), dS = P * (dP - delta) * scale   -> bf16 [query][key] tiles in smem
//     dV += P^T dO, dK += dS^T Q               (A operands read MN-major from those tiles; TMEM [256,320), [320,384))
//     dQ_m = dS K -> TMEM [384,448) / [448,512) by tile parity -> fp32 RED into the dq accumulator
//   Software pipeline: after the P/dS tile of query tile m is in smem, one thread issues the three accumulating
//   GEMMs of tile m AND S/dP of tile m+1 behind them, one commit for the lot; while the tensor pipe works the
//   512 threads drain dQ of tile m-1 (TMEM -> RED.v4), so the reduction traffic and the MMAs overlap and the
//   next tile's scores are ready when the threads come back.
//     d_rel[j-i] += dS: per-warp diagonal sums by lane shuffles, accumulated in smem, one global atomic per offset
// =================================================================================
constexpr int kBwdComputeWarps = 16;
constexpr int kBwdComputeThreads = kBwdComputeWarps * 32;
constexpr int kBwdThreads2 = kBwdComputeThreads + 32;     // + the issuer warp
constexpr uint32_t kBwdTmemCols = 512;   // S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ [384,448), [448,512)
constexpr int kBwdStages = 2;            // Q / dO ring
constexpr int kBwdStageBytes = 2 * kBM * kD * 2;
constexpr int kBOffK = 0;
constexpr int kBOffV = kBOffK + kBN * kD * 2;             // 16 KB
constexpr int kBOffQ = kBOffV + kBN * kD * 2;             // 32 KB: 2 stages x (Q 16 KB, dO 16 KB)
constexpr int kBOffP = kBOffQ + kBwdStages * kBwdStageBytes;   // 96 KB
constexpr int kBOffdS = kBOffP + kBM * kBN * 2;           // 128 KB
constexpr int kBOffStg = kBOffdS + kBM * kBN * 2;         // 160 KB: fp32 dQ staging, two [128][32] SW128 halves
constexpr int kBOffBar = kBOffStg + kBM * kD * 4;         // 192 KB
constexpr int kBOffFloats = kBOffBar + 64;

struct AttnBwdParams {
  const float* lse;
  const float* delta;         // (B,H,Sq)
  const float* rel_bias;
  const float* key_add;
  float* dq_accum;            // (B,Sq,H,64) fp32, zero-initialised
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  float* d_rel;               // (H, Sq+Sk-1) fp32 accumulated, or null
  int B, H, Sq, Sk;
  long long dk_stride_b, dk_stride_s, dk_stride_h;
  long long dv_stride_b, dv_stride_s, dv_stride_h;
  float scale;
  int causal;
  uint32_t drop_thr8;
  float drop_scale;
  uint64_t seed, offset;
  const unsigned long long* rng_base;
  const uint8_t* scp_bucket;  // SaL SCP bias (see AttnFwdParams)
  const float* scp_tab;
  float* d_scp;               // (H, 32) fp32 accumulated, or null
  int scp_q0, scp_L;
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

// FULL = false compiles the SaL spatial-bias code and the causal test out.  Both sit in the per-element loops as
// run-time flags, so every score of a plain bidirectional launch pays their predicated-off bucket extraction, table
// load and add, and a compare + select for the diagonal.  The launcher uses that variant for non-causal launches
// without an SCP bias only when PVQA_ATTN_BWD_LEAN=1 (opt-in until it has been validated on a device; FULL = true is,
// instruction for instruction, the kernel that was validated).
template <bool HAS_REL, bool DROP, bool FULL = true>
__global__ void __launch_bounds__(kBwdThreads2, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an OFFSET into the __shared__ array so that the compiler
  // keeps the shared address space (32-bit LDS/STS instead of generic 64-bit LD/ST for every smem access)
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kBOffBar);
  uint64_t* bar_ld = bar_kv + 1;          // [kBwdStages] Q/dO ring
  uint64_t* bar_s = bar_kv + 3;           // S/dP of a query tile are in TMEM
  uint64_t* bar_g = bar_kv + 4;           // the three accumulating GEMMs of a query tile are done
  uint64_t* bar_stg = bar_kv + 5;         // the dQ staging tile has been read by its reduce
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 6);
  const int n_rel = p.Sq + p.Sk - 1;
  const int n_win = p.Sq + kBN - 1;                              // rel offsets this key tile can see
  float* s_kadd = reinterpret_cast<float*>(smem + kBOffFloats);  // [kBN], -inf beyond Sk
  float* s_rel = s_kadd + kBN;                                   // [kRelPad + n_win] bias * log2e (index w + kRelPad)
  float* s_drel = s_rel + kRelPad + n_win;                       // [n_win] gradient accumulator (smem atomics)
  float* s_scp = s_drel + n_win;                                 // [32] SCP table of this head
  float* s_dscp = s_scp + 32;                                    // [16 warps][32] SCP gradient bins
  const bool has_scp = FULL && HAS_REL && p.scp_bucket != nullptr;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == kBwdComputeWarps;       // warp 16: TMA, tcgen05.mma and the dQ reduce, nothing else
  const int rowl = (warp & 3) * 32 + lane;     // row inside the 128-row tile == TMEM lane
  const int qd = warp >> 2;                    // which quarter of the columns this thread owns (compute warps)
  const int j0 = blockIdx.x * kBN;
  const int h = blockIdx.y, b = blockIdx.z;
  PVQA_TRACE(0);

  const int m_tiles = (p.Sq + kBM - 1) / kBM;
  const int m_first = p.causal ? (j0 / kBM) : 0;      // query tiles entirely above the diagonal see nothing
  const int n_it = m_tiles > m_first ? m_tiles - m_first : 0;
  if (is_issuer) {
    if (lane == 0) {
      tc05::prefetch_tmap(&tmQ); tc05::prefetch_tmap(&tmK); tc05::prefetch_tmap(&tmV); tc05::prefetch_tmap(&tmdO);
      tc05::prefetch_tmap(&tmdQ);
      tc05::mbar_init(bar_kv, 1);
      for (int st = 0; st < kBwdStages; ++st) tc05::mbar_init(bar_ld + st, 1);
      tc05::mbar_init(bar_s, 1); tc05::mbar_init(bar_g, 1); tc05::mbar_init(bar_stg, 1);
      tc05::fence_barrier_init();
      // the loads go out before anything else: K, V and the first two query tiles
      tc05::mbar_expect_tx(bar_kv, 2 * kBN * kD * 2);
      tc05::tma_load_4d(smem + kBOffK, &tmK, bar_kv, 0, h, j0, b);
      tc05::tma_load_4d(smem + kBOffV, &tmV, bar_kv, 0, h, j0, b);
      for (int st = 0; st < kBwdStages && st < n_it; ++st) {
        uint8_t* dst = smem + kBOffQ + st * kBwdStageBytes;
        tc05::mbar_expect_tx(bar_ld + st, kBwdStageBytes);
        tc05::tma_load_4d(dst, &tmQ, bar_ld + st, 0, h, (m_first + st) * kBM, b);
        tc05::tma_load_4d(dst + kBM * kD * 2, &tmdO, bar_ld + st, 0, h, (m_first + st) * kBM, b);
      }
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_slot, kBwdTmemCols);
    tc05::tmem_relinquish();
  } else {
    if (tid < kBN) {
      const int j = j0 + tid;
      s_kadd[tid] = (j < p.Sk) ? (p.key_add ? p.key_add[(long long)b * p.Sk + j] * kLog2e : 0.f) : -INFINITY;
    }
    if (HAS_REL) {
      // window of relative offsets: global rel index r = j - i + Sq - 1 = j0 + w, w in [0, n_win)
      for (int x = tid; x < kRelPad + n_win; x += kBwdComputeThreads) {
        const int r = j0 + x - kRelPad;
        s_rel[x] = (x >= kRelPad && r < n_rel) ? p.rel_bias[(long long)h * n_rel + r] * kLog2e : 0.f;
      }
      for (int x = tid; x < n_win; x += kBwdComputeThreads) s_drel[x] = 0.f;
      if (has_scp) {
        if (tid < 32) s_scp[tid] = p.scp_tab[h * 32 + tid] * kLog2e;
        s_dscp[tid] = 0.f;                       // 512 threads == 16 x 32 bins
      }
    }
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  PVQA_TRACE(1);

  if (is_issuer) {
    // ================= issuer warp =================
    const uint32_t idesc_s = tc05::idesc_bf16(kBM, kBN, 0, 0);
    const uint32_t idesc_dkv = tc05::idesc_bf16(kBN, kD, 1, 1);   // A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
    const uint32_t idesc_dq = tc05::idesc_bf16(kBM, kD, 0, 1);    // A = dS (K-major), B = K (MN-major)
    const uint32_t k_addr = tc05::smem_u32(smem + kBOffK), v_addr = tc05::smem_u32(smem + kBOffV);
    const uint32_t p_addr = tc05::smem_u32(smem + kBOffP), ds_addr = tc05::smem_u32(smem + kBOffdS);
    const uint32_t q_base = tc05::smem_u32(smem + kBOffQ);
    auto issue_s_dp = [&](int st) {           // S = Q K^T and dP = dO V^T of the query tile in ring stage st
      const uint32_t qa = q_base + st * kBwdStageBytes, da = qa + kBM * kD * 2;
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks)
        tc05::mma_bf16_ss(tmem_base, tc05::smem_desc_sw128(qa + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(k_addr + ks * 32, 16, 1024), idesc_s, ks > 0);
#pragma unroll
      for (int ks = 0; ks < kD / 16; ++ks)
        tc05::mma_bf16_ss(tmem_base + kBN, tc05::smem_desc_sw128(da + ks * 32, 16, 1024),
                          tc05::smem_desc_sw128(v_addr + ks * 32, 16, 1024), idesc_s, ks > 0);
    };
    auto reduce_dq = [&](int itp) {           // staging tile (fp32, two [128][32] SW128 halves) += into dq_accum
      tc05::tma_reduce_add_4d(&tmdQ, smem + kBOffStg, 0, h, (m_first + itp) * kBM, b);
      tc05::tma_reduce_add_4d(&tmdQ, smem + kBOffStg + kBM * 128, 32, h, (m_first + itp) * kBM, b);
      tc05::bulk_commit_group();
    };
    if (lane == 0 && n_it > 0) {
      tc05::mbar_wait(bar_kv, 0);
      tc05::mbar_wait(bar_ld, 0);
      tc05::tc_fence_after_sync();
      issue_s_dp(0);
      tc05::mma_commit(bar_s);
    }
    for (int it = 0; it < n_it; ++it) {
      // (1) every compute thread holds S/dP of tile it in registers: TMEM S/dP can take the next tile's scores
      tc05::named_bar_sync(1, kBwdThreads2);
      if (lane == 0 && it + 1 < n_it) {
        tc05::tc_fence_after_sync();
        const int st1 = (it + 1) % kBwdStages;
        tc05::mbar_wait(bar_ld + st1, ((it + 1) / kBwdStages) & 1);
        tc05::tc_fence_after_sync();
        issue_s_dp(st1);
        tc05::mma_commit(bar_s);
      }
      __syncwarp();
      // (2) P/dS of tile it are in smem and dQ(it-1) is staged
      tc05::named_bar_sync(3, kBwdThreads2);
      if (lane == 0) {
        tc05::tc_fence_after_sync();
        const int stg = it % kBwdStages;
        const uint32_t q_addr = q_base + stg * kBwdStageBytes, do_addr = q_addr + kBM * kD * 2;
        if (it > 0) reduce_dq(it - 1);
        const uint32_t dq_col = tmem_base + 384 + (it & 1) * 64;
#pragma unroll
        for (int ks = 0; ks < kBM / 16; ++ks)     // dV += P^T dO   (K = 128 query rows, 16 per step)
          tc05::mma_bf16_ss(tmem_base + 256, tc05::smem_desc_sw128(p_addr + ks * 2048, kBM * 128, 1024),
                            tc05::smem_desc_sw128(do_addr + ks * 2048, 16, 1024), idesc_dkv, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < kBM / 16; ++ks)     // dK += dS^T Q
          tc05::mma_bf16_ss(tmem_base + 320, tc05::smem_desc_sw128(ds_addr + ks * 2048, kBM * 128, 1024),
                            tc05::smem_desc_sw128(q_addr + ks * 2048, 16, 1024), idesc_dkv, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ++ks)     // dQ_m = dS K    (K = 128 keys)
          tc05::mma_bf16_ss(dq_col,
                            tc05::smem_desc_sw128(ds_addr + (ks >> 2) * (kBM * 128) + (ks & 3) * 32, 16, 1024),
                            tc05::smem_desc_sw128(k_addr + ks * 2048, 16, 1024), idesc_dq, ks > 0);
        tc05::mma_commit(bar_g);
        PVQA_TRACE(3 + 5 * it);
        if (it > 0) {                             // the reduce has read the staging tile: hand it back
          tc05::bulk_wait_group_read0();
          tc05::mbar_arrive(bar_stg);
        }
        if (it + kBwdStages < n_it) {             // this tile's ring stage is free once its GEMMs are done
          tc05::mbar_wait(bar_g, it & 1);
          uint8_t* dst = smem + kBOffQ + stg * kBwdStageBytes;
          const int i_next = (m_first + it + kBwdStages) * kBM;
          tc05::mbar_expect_tx(bar_ld + stg, kBwdStageBytes);
          tc05::tma_load_4d(dst, &tmQ, bar_ld + stg, 0, h, i_next, b);
          tc05::tma_load_4d(dst + kBM * kD * 2, &tmdO, bar_ld + stg, 0, h, i_next, b);
        }
        PVQA_TRACE(4 + 5 * it);
      }
      __syncwarp();
    }
    if (n_it > 0) {
      tc05::named_bar_sync(3, kBwdThreads2);      // dQ of the last tile is staged
      if (lane == 0) {
        reduce_dq(n_it - 1);
        tc05::bulk_wait_group0();                 // all reductions performed before the CTA retires
      }
      __syncwarp();
    } else if (lane == 0) {
      tc05::mbar_wait(bar_kv, 0);                 // never leave with a TMA write in flight
    }
  } else {
    // ================= 16 compute warps =================
    const float sl2 = p.scale * kLog2e;
    const uint32_t thr4 = p.drop_thr8 * 0x01010101u;
    const int jl0 = qd * 32;                      // first local key column of this thread
    const uint64_t rng_off = p.offset + ((DROP && p.rng_base) ? *p.rng_base : 0ull);
    uint8_t* stg_row = smem + kBOffStg + (qd >> 1) * (kBM * 128) + rowl * 128;
    // dQ of query tile itp (complete in TMEM) -> fp32 staging tile in smem (the issuer reduces it into dq_accum)
    auto stage_dq = [&](int itp) {
      if (itp > 0) tc05::mbar_wait(bar_stg, (itp - 1) & 1);       // reduce of tile itp-1 has read the buffer
      uint32_t r[16];
      tc05::tmem_ld_32x16(tmem_row + 384 + (itp & 1) * 64 + qd * 16, r);
      tc05::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = ((qd & 1) * 4 + q) ^ (rowl & 7);
        *reinterpret_cast<uint4*>(stg_row + chunk * 16) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
      }
    };
    // per-row statistics of a query tile, fetched one tile ahead so the global-load latency hides behind the math
    auto load_stats = [&](int itn, float& l_out, float& d_out) {
      const int in = (m_first + itn) * kBM + rowl;
      l_out = -INFINITY; d_out = 0.f;
      if (itn < n_it && in < p.Sq) {
        const long long ri = ((long long)b * p.H + h) * p.Sq + in;
        l_out = p.lse[ri];
        d_out = p.delta[ri];
      }
    };
    float lse_nx, delta_nx;
    load_stats(0, lse_nx, delta_nx);
    for (int it = 0; it < n_it; ++it) {
      const int i0 = (m_first + it) * kBM;
      const int i = i0 + rowl;
      const bool row_ok = i < p.Sq;
      // +inf => p = exp2(s - inf) = 0 for dead rows and for rows whose softmax was empty (lse = -inf)
      const float lse2 = (lse_nx != -INFINITY) ? lse_nx * kLog2e : INFINITY;
      const float delta = delta_nx;
      load_stats(it + 1, lse_nx, delta_nx);
      const bool diag = FULL && p.causal && (j0 + kBN - 1 > i0);
      const float* relrow = s_rel + (p.Sq - 1 - i) + kRelPad;      // relrow[jl] = bias of local key jl for this row
      // P/dS smem is still read by the GEMMs of tile it-1: wait for them right before the stores (long done by then)
      auto wait_prev_gemms = [&]() {
        if (it > 0) {
          tc05::mbar_wait(bar_g, (it - 1) & 1);
          tc05::tc_fence_after_sync();
        }
      };
      tc05::mbar_wait(bar_s, it & 1);
      tc05::tc_fence_after_sync();
      PVQA_TRACE(2 + 5 * it);

      // ---- P and dS for this thread's 32 columns ----
      const bool dead = j0 + jl0 >= p.Sk || i0 + (warp & 3) * 32 >= p.Sq;   // warp-uniform: keys past Sk / rows past Sq
      uint8_t* prow = smem + kBOffP + (qd >> 1) * (kBM * 128) + rowl * 128;
      uint8_t* dsrow = smem + kBOffdS + (qd >> 1) * (kBM * 128) + rowl * 128;
      if (dead) {
        tc05::tc_fence_before_sync();
        tc05::named_bar_arrive(1, kBwdThreads2);
        wait_prev_gemms();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((qd & 1) * 4 + q) ^ (rowl & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dsrow + chunk * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        uint32_t rs[32], rp[32];
        tc05::tmem_ld_32x32(tmem_row + jl0, rs);
        tc05::tmem_ld_32x32(tmem_row + kBN + jl0, rp);
        tc05::tmem_ld_wait();
        // S/dP of this tile now live in registers: the issuer may overwrite TMEM with the next tile's
        tc05::tc_fence_before_sync();
        tc05::named_bar_arrive(1, kBwdThreads2);
        // SCP bucket ids of the 32 columns (two 16-byte halves, each inside or outside the OCR block)
        const int jj0 = j0 + jl0 - p.scp_q0;
        uint32_t bkw[8];
        bool scp_h[2] = {false, false};
        if (has_scp) {
          const bool row_in = row_ok && i >= p.scp_q0 && i < p.scp_q0 + p.scp_L;
          const uint8_t* scp_row = p.scp_bucket + ((long long)b * p.scp_L + (i - p.scp_q0)) * p.scp_L;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int jj = jj0 + hf * 16;
            scp_h[hf] = row_in && jj >= 0 && jj < p.scp_L;
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (scp_h[hf]) u = __ldg(reinterpret_cast<const uint4*>(scp_row + jj));
            bkw[hf * 4] = u.x; bkw[hf * 4 + 1] = u.y; bkw[hf * 4 + 2] = u.z; bkw[hf * 4 + 3] = u.w;
          }
        }
        // dropout: the 1/keep factor rides in the exponent (pv = P/keep), so P_drop = pv & mask and
        // dS = P (M/keep dP - delta) scale = pv (M dP - keep delta) scale = pv * fma(M*scale, dP, -keep*delta*scale)
        const float lse_k = DROP ? lse2 - log2f(p.drop_scale) : lse2;
        const float nds = -(DROP ? delta / p.drop_scale : delta) * p.scale;
        float pv[32], dsv[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          // key term: one 16-byte broadcast load per 4 columns (every lane reads the same address)
          const float4 ka4 = *reinterpret_cast<const float4*>(s_kadd + jl0 + (x & ~3));
          float bias = (x & 3) == 0 ? ka4.x : (x & 3) == 1 ? ka4.y : (x & 3) == 2 ? ka4.z : ka4.w;
          if (HAS_REL) bias += relrow[jl0 + x];
          if (has_scp && scp_h[x >> 4]) bias += s_scp[(bkw[x >> 2] >> (8 * (x & 3))) & 31u];
          float sc = fmaf(__uint_as_float(rs[x]), sl2, bias);
          if (diag && (j0 + jl0 + x > i)) sc = -INFINITY;
          pv[x] = fast_exp2(sc - lse_k);
        }
        if (DROP) {
          const uint64_t drop_row = ((uint64_t)(b * p.H + h) * p.Sq + min(i, p.Sq - 1)) * (uint64_t)((p.Sk + 15) >> 4);
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            const uint4 kb = keep_bytes16(p.seed, rng_off, drop_row + ((j0 + jl0) >> 4) + g2, thr4);
            const uint32_t kw[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
            for (int xx = 0; xx < 16; ++xx) {
              const int x = g2 * 16 + xx;
              const uint32_t m = PVQA_BYTE_MASK(kw[xx >> 2], xx & 3);
              dsv[x] = pv[x] * fmaf(__uint_as_float(__float_as_uint(p.scale) & m), __uint_as_float(rp[x]), nds);
              pv[x] = __uint_as_float(__float_as_uint(pv[x]) & m);
            }
          }
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) dsv[x] = pv[x] * fmaf(p.scale, __uint_as_float(rp[x]), nds);
        }
        if (has_scp && p.d_scp) {
          // d_scp[bucket] += dS on the OCR x OCR block: per-warp shared-memory bins (conflicting lanes serialise)
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (scp_h[x >> 4]) atomicAdd(s_dscp + warp * 32 + ((bkw[x >> 2] >> (8 * (x & 3))) & 31u), dsv[x]);
        }
        if (HAS_REL && p.d_rel) {
          // d_rel[j-i] += dS: this warp holds a 32x32 block (lane = row, register = column).  Lane L collects the
          // diagonals d' = col - row == L (mod 32): one shuffle per column, two accumulators for the wrap.
          float accp = 0.f, accn = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float vsh = __shfl_sync(0xffffffffu, dsv[c], c - lane);   // source lane taken mod 32
            if (lane <= c) accp += vsh; else accn += vsh;
          }
          // window index of rel = (j0+jl) - (i0+il) + Sq-1, minus j0;  jl - il = (jl0 - 32*(warp&3)) + d'
          const int wpos = jl0 - (warp & 3) * 32 + lane - i0 + p.Sq - 1;
          if (wpos >= 0 && wpos < n_win) atomicAdd(s_drel + wpos, accp);
          if (lane > 0 && wpos - 32 >= 0 && wpos - 32 < n_win) atomicAdd(s_drel + wpos - 32, accn);
        }
        wait_prev_gemms();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((qd & 1) * 4 + q) ^ (rowl & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = pack8(pv + q * 8);
          *reinterpret_cast<uint4*>(dsrow + chunk * 16) = pack8(dsv + q * 8);
        }
      }
      PVQA_TRACE(3 + 5 * it);
      if (it > 0) stage_dq(it - 1);               // complete since bar_g(it-1)
      tc05::fence_proxy_async_smem();
      tc05::tc_fence_before_sync();
      tc05::named_bar_arrive(3, kBwdThreads2);
      PVQA_TRACE(4 + 5 * it);
    }
    if (n_it > 0) {
      tc05::mbar_wait(bar_g, (n_it - 1) & 1);
      tc05::tc_fence_after_sync();
      PVQA_TRACE(25);
      stage_dq(n_it - 1);
      tc05::fence_proxy_async_smem();
      tc05::tc_fence_before_sync();
      tc05::named_bar_arrive(3, kBwdThreads2);
      PVQA_TRACE(26);
    }
    // ---- epilogue: dV, dK rows (key j0 + rowl), columns [16*qd, +16) ----
    {
      const int j = j0 + rowl;
      uint32_t rv[16], rk[16];
      if (n_it > 0) {
        tc05::tmem_ld_32x16(tmem_row + 256 + qd * 16, rv);
        tc05::tmem_ld_32x16(tmem_row + 320 + qd * 16, rk);
        tc05::tmem_ld_wait();
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) { rv[x] = 0u; rk[x] = 0u; }
      }
      if (j < p.Sk) {
        __nv_bfloat16* dvrow = p.dv + b * p.dv_stride_b + j * p.dv_stride_s + h * p.dv_stride_h + qd * 16;
        __nv_bfloat16* dkrow = p.dk + b * p.dk_stride_b + j * p.dk_stride_s + h * p.dk_stride_h + qd * 16;
        float fv[16], fk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) { fv[x] = __uint_as_float(rv[x]); fk[x] = __uint_as_float(rk[x]); }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          *reinterpret_cast<uint4*>(dvrow + c * 8) = pack8(fv + c * 8);
          *reinterpret_cast<uint4*>(dkrow + c * 8) = pack8(fk + c * 8);
        }
      }
    }
    PVQA_TRACE(27);
    tc05::named_bar_sync(2, kBwdComputeThreads);       // every warp's smem bins are final
    PVQA_TRACE(28);
    if (HAS_REL && p.d_rel) {
      const float inv_scale = 1.0f / p.scale;          // the smem tile holds scale * dS
      for (int w = tid; w < n_win; w += kBwdComputeThreads) {
        const int r = j0 + w;
        const float g = s_drel[w];
        if (r < n_rel && g != 0.f) atomicAdd(p.d_rel + (long long)h * n_rel + r, g * inv_scale);
      }
    }
  }
  PVQA_TRACE(30);
  tc05::tc_fence_before_sync();
  __syncthreads();
  PVQA_TRACE(31);
  if (has_scp && p.d_scp && tid < 32) {
    float g = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdComputeWarps; ++w) g += s_dscp[w * 32 + tid];
    if (g != 0.f) atomicAdd(p.d_scp + h * 32 + tid, g / p.scale);
  }
  if (is_issuer) tc05::tmem_dealloc(tmem_base, kBwdTmemCols);
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

#ifdef PVQA_ATTN_TRACE
extern "C" int pvqa_debug_attn_trace(long long* host, int clear) {
  if (host && cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * 64 * 64) != cudaSuccess) return PVQA_ERR_CUDA;
  if (clear) {
    static long long zeros[64 * 64];
    if (cudaMemcpyToSymbol(g_attn_trace, zeros, sizeof(zeros)) != cudaSuccess) return PVQA_ERR_CUDA;
  }
  return PVQA_OK;
}
#endif

extern "C" int pvqa_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                             const float* rel_bias, const float* key_add, int64_t B, int64_t H, int64_t Sq,
                             int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                             int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t v_stride_b,
                             int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b, int64_t o_stride_s,
                             int64_t o_stride_h, float scale, int causal, float dropout_p, uint64_t seed,
                             uint64_t offset, const uint8_t* scp_bucket, const float* scp_table, int64_t scp_q0,
                             int64_t scp_L, void* stream) {
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "attn_fwd: dropout_p must be in [0,1)");
  if (scp_bucket) {
    PVQA_REQUIRE(rel_bias && scp_table, PVQA_ERR_NULL, "attn_fwd: the SCP bias needs rel_bias and scp_table");
    PVQA_REQUIRE(!causal && Sq == Sk && scp_q0 >= 0 && scp_L > 0 && scp_q0 + scp_L <= Sk, PVQA_ERR_SHAPE,
                 "attn_fwd: bad SCP block");
    PVQA_REQUIRE(scp_q0 % 16 == 0 && scp_L % 16 == 0 && aligned16(scp_bucket), PVQA_ERR_ALIGN,
                 "attn_fwd: SCP block offset/size must be multiples of 16");
  }
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
  const int64_t n_kpad = (Sk + kBN - 1) / kBN * kBN;
  const int64_t n_floats = n_kpad + (rel_bias ? kRelPad + Sq + n_kpad + 32 : 0);
  const size_t smem_bytes = 1024 + kOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 113 * 1024, PVQA_ERR_SHAPE,
               "attn_fwd: Sq/Sk too large for the bias staging buffers (2 CTAs per SM budget)");
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
  auto kern = rel ? (drop ? attn_fwd_kernel<true, true> : attn_fwd_kernel<true, false>)
                  : (drop ? attn_fwd_kernel<false, true> : attn_fwd_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  const int vi = (rel ? 2 : 0) + (drop ? 1 : 0);
  if (!attr_set[vi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[vi] = true;
  }
  dim3 grid((unsigned)((Sq + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  kern<<<grid, kFwdThreads, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                             const float* lse, const float* rel_bias, const float* key_add, float* dq_accum,
                             void* dk, void* dv, float* d_rel_bias, float* delta_ws, int64_t B, int64_t H,
                             int64_t Sq, int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s,
                             int64_t q_stride_h, int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                             int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b,
                             int64_t o_stride_s, int64_t o_stride_h, int64_t do_stride_b, int64_t do_stride_s,
                             int64_t do_stride_h, int64_t dk_stride_b, int64_t dk_stride_s, int64_t dk_stride_h,
                             int64_t dv_stride_b, int64_t dv_stride_s, int64_t dv_stride_h, float scale, int causal,
                             float dropout_p, uint64_t seed, uint64_t offset, const uint8_t* scp_bucket,
                             const float* scp_table, float* d_scp_table, int64_t scp_q0, int64_t scp_L, void* stream) {
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "attn_bwd: dropout_p must be in [0,1)");
  if (scp_bucket) {
    PVQA_REQUIRE(rel_bias && scp_table, PVQA_ERR_NULL, "attn_bwd: the SCP bias needs rel_bias and scp_table");
    PVQA_REQUIRE(!causal && Sq == Sk && scp_q0 >= 0 && scp_L > 0 && scp_q0 + scp_L <= Sk, PVQA_ERR_SHAPE,
                 "attn_bwd: bad SCP block");
    PVQA_REQUIRE(scp_q0 % 16 == 0 && scp_L % 16 == 0 && aligned16(scp_bucket), PVQA_ERR_ALIGN,
                 "attn_bwd: SCP block offset/size must be multiples of 16");
  }
  PVQA_REQUIRE(D == kD, PVQA_ERR_SHAPE, "attn_bwd: head dim %lld unsupported (kernel is specialised for 64)", (long long)D);
  PVQA_REQUIRE(B >= 0 && H > 0 && Sq >= 0 && Sk >= 0, PVQA_ERR_SHAPE, "attn_bwd: bad dimension");
  if (B == 0 || Sq == 0 || Sk == 0) return PVQA_OK;
  PVQA_REQUIRE(q && k && v && o && d_o && lse && dq_accum && dk && dv && delta_ws, PVQA_ERR_NULL, "attn_bwd: NULL pointer");
  PVQA_REQUIRE(!causal || Sq == Sk, PVQA_ERR_SHAPE, "attn_bwd: causal requires Sq == Sk");
  PVQA_REQUIRE(scale != 0.f, PVQA_ERR_SHAPE, "attn_bwd: scale must be non-zero");
  PVQA_REQUIRE(!d_rel_bias || rel_bias, PVQA_ERR_NULL, "attn_bwd: d_rel_bias requested without rel_bias");
  PVQA_REQUIRE(H <= 65535 && B <= 65535, PVQA_ERR_SHAPE, "attn_bwd: H and B must be <= 65535");
  auto al8 = [](int64_t a, int64_t b2, int64_t c) { return a % 8 == 0 && b2 % 8 == 0 && c % 8 == 0; };
  PVQA_REQUIRE(aligned16(o) && aligned16(d_o) && aligned16(dk) && aligned16(dv) && aligned16(dq_accum) &&
                   al8(o_stride_b, o_stride_s, o_stride_h) && al8(do_stride_b, do_stride_s, do_stride_h) &&
                   al8(dk_stride_b, dk_stride_s, dk_stride_h) && al8(dv_stride_b, dv_stride_s, dv_stride_h),
               PVQA_ERR_ALIGN, "attn_bwd: rows must be 16-byte aligned");
  const int64_t n_win = Sq + kBN - 1;
  const int64_t n_floats = kBN + (rel_bias ? kRelPad + 2 * n_win + 32 + 512 : 0);
  const size_t smem_bytes = 1024 + kBOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 225 * 1024, PVQA_ERR_SHAPE, "attn_bwd: Sq too large for the bias window buffers");
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap(&tq, q, B, Sq, H, q_stride_b, q_stride_s, q_stride_h, kBM, "q"))) return rc;
  if ((rc = make_tmap(&tk, k, B, Sk, H, k_stride_b, k_stride_s, k_stride_h, kBN, "k"))) return rc;
  if ((rc = make_tmap(&tv, v, B, Sk, H, v_stride_b, v_stride_s, v_stride_h, kBN, "v"))) return rc;
  if ((rc = make_tmap(&tdo, d_o, B, Sq, H, do_stride_b, do_stride_s, do_stride_h, kBM, "d_o"))) return rc;
  CUtensorMap tdq;          // fp32 (B,Sq,H,64) contiguous accumulator, reduced into by 32-column half tiles
  {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(PVQA_ERR_CUDA, "attn_bwd: cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[4] = {(cuuint64_t)kD, (cuuint64_t)H, (cuuint64_t)Sq, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)kD * 4, (cuuint64_t)H * kD * 4, (cuuint64_t)Sq * H * kD * 4};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)kBM, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_accum, gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVQA_ERR_CUDA, "attn_bwd: cuTensorMapEncodeTiled(dq_accum) failed with %d", (int)r);
  }
  cudaStream_t st = (cudaStream_t)stream;
  AttnPrepParams pp{};
  pp.o = reinterpret_cast<const __nv_bfloat16*>(o); pp.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
  pp.delta = delta_ws; pp.B = (int)B; pp.H = (int)H; pp.Sq = (int)Sq;
  pp.o_stride_b = o_stride_b; pp.o_stride_s = o_stride_s; pp.o_stride_h = o_stride_h;
  pp.do_stride_b = do_stride_b; pp.do_stride_s = do_stride_s; pp.do_stride_h = do_stride_h;
  {
    const long long n = (long long)B * H * Sq * 8;         // 8 lanes per row
    long long need = (n + 255) / 256, cap = (long long)num_sms() * 8;
    attn_bwd_prep_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(pp);
    count_launch();
    PVQA_CHECK_LAUNCH("attn_bwd(prep)");
  }
  AttnBwdParams p{};
  p.lse = lse; p.delta = delta_ws; p.rel_bias = rel_bias; p.key_add = key_add; p.dq_accum = dq_accum;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.d_rel = d_rel_bias;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.dk_stride_b = dk_stride_b; p.dk_stride_s = dk_stride_s; p.dk_stride_h = dk_stride_h;
  p.dv_stride_b = dv_stride_b; p.dv_stride_s = dv_stride_s; p.dv_stride_h = dv_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset; p.rng_base = g_rng_base;
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.d_scp = d_scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  const bool rel = rel_bias != nullptr, drop = p.drop_thr8 != 0;
  auto kern = rel ? (drop ? attn_bwd_kernel<true, true> : attn_bwd_kernel<true, false>)
                  : (drop ? attn_bwd_kernel<false, true> : attn_bwd_kernel<false, false>);
  const char* lean_env = getenv("PVQA_ATTN_BWD_LEAN");      // read per call: cheap, and a process can compare both variants
  const bool lean_opt_in = lean_env && lean_env[0] == '1';
  const bool lean = lean_opt_in && scp_bucket == nullptr && !causal;       // encoder self- and decoder cross-attention
  if (lean)
    kern = rel ? (drop ? attn_bwd_kernel<true, true, false> : attn_bwd_kernel<true, false, false>)
               : (drop ? attn_bwd_kernel<false, true, false> : attn_bwd_kernel<false, false, false>);
  static bool attr_set[8] = {false, false, false, false, false, false, false, false};
  const int vi = (lean ? 4 : 0) + (rel ? 2 : 0) + (drop ? 1 : 0);
  if (!attr_set[vi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[vi] = true;
  }
  dim3 grid((unsigned)((Sk + kBN - 1) / kBN), (unsigned)H, (unsigned)B);
  kern<<<grid, kBwdThreads2, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_bwd");
  return PVQA_OK;
}

#include "attn_fwd2.cuh"
#include "attn_fwd3.cuh"
