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
                             int64_t o_stride_h, float scale, int causal, void* stream) {
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
  dim3 grid((unsigned)((Sq + kBM - 1) / kBM), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<<<grid, kAttnThreads, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_fwd");
  return PVQA_OK;
}
