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

}  // namespace pvqa

#include "attn_fwd.cuh"

#include "attn_bwd.cuh"

namespace pvqa {

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

// dropout parameters of a launch (AttnDrop, attn_fwd.cuh); returns the quantised threshold (0 = dropout off)
static uint32_t fill_drop(AttnDrop* d, float dropout_p, uint64_t seed, uint64_t offset, int64_t Sk) {
  const uint32_t thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  for (int r = 0; r < 7; ++r) {
    d->rk[r][0] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
    d->rk[r][1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  for (int k = 0; k < 8; ++k) d->tmask[k] = ((thr8 >> k) & 1u) ? 0xffffffffu : 0u;
  d->keep_scale = thr8 ? 256.f / (256.f - (float)thr8) : 1.f;
  d->m_shift = log2f(d->keep_scale);
  d->blk_per_row = (uint32_t)((Sk + 31) / 32);
  d->offset = offset;
  d->rng_base = g_rng_base;
  return thr8;
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
  PVQA_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0 && o_stride_s % 8 == 0 && o_stride_h % 8 == 0 &&
                   o_stride_b % 8 == 0,
               PVQA_ERR_ALIGN, "attn_fwd: output rows must be 16-byte aligned");
  // key tiles: n_full of width 128, then the remainder r as ONE tile of width 32, 64 or 128 (masked past Sk).  A tile
  // costs a fixed ~1.5 k cycles of barriers / TMEM traffic whatever its width, so covering r = 71 (S = 327) with a 64-
  // and a 32-wide tile — less padding, one tile more — measured slower than one masked 128-wide tile: 153 vs 139 us
  // per encoder launch, 67 vs 58 us at S = 197 (gpu call 41).  The kernel still accepts a second remainder tile (w_b).
  const int n_full = (int)(Sk / kBN), rem = (int)(Sk % kBN);
  const int w_a = rem == 0 ? 0 : rem <= 32 ? 32 : rem <= 64 ? 64 : 128;
  const int w_b = 0;
  const int n_kt = n_full + (w_a ? 1 : 0) + (w_b ? 1 : 0);
  const int n_kpad = n_kt * kBN;
  const int64_t n_floats = 2 * (int64_t)n_kpad + (rel_bias ? 2 * (int64_t)f_rel_copy_stride(n_kpad) + 32 : 0);
  const size_t smem_bytes = 1024 + kFOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 227 * 1024, PVQA_ERR_SHAPE, "attn_fwd: Sk too large for the bias staging buffers");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap(&tq, q, B, Sq, H, q_stride_b, q_stride_s, q_stride_h, kBM, "q"))) return rc;
  if ((rc = make_tmap(&tk, k, B, Sk, H, k_stride_b, k_stride_s, k_stride_h, kBN, "k"))) return rc;
  if ((rc = make_tmap(&tv, v, B, Sk, H, v_stride_b, v_stride_s, v_stride_h, kBN, "v"))) return rc;
  AttnFwdParams p{};
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.lse = lse; p.rel_bias = rel_bias; p.key_add = key_add;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.o_stride_b = o_stride_b; p.o_stride_s = o_stride_s; p.o_stride_h = o_stride_h;
  p.sl2 = scale * kLog2e;
  const uint32_t thr8 = fill_drop(&p.drop, dropout_p, seed, offset, Sk);
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  p.n_qt = (int)((Sq + kBM - 1) / kBM); p.n_kt = n_kt; p.n_full = n_full; p.w_a = w_a; p.w_b = w_b;
  const int64_t n_items = H * (int64_t)p.n_qt * B;
  PVQA_REQUIRE(n_items < (1ll << 30), PVQA_ERR_SHAPE, "attn_fwd: too many (batch, head, query tile) items");
  p.n_items = (int)n_items;
  const bool rel = rel_bias != nullptr, drop = thr8 != 0, scp = scp_bucket != nullptr, cz = causal != 0;
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnFwdParams);
  static const Kern table[10] = {
      attn_fwd_kernel<false, false, false, false>, attn_fwd_kernel<false, true, false, false>,
      attn_fwd_kernel<false, false, true, false>,  attn_fwd_kernel<false, true, true, false>,
      attn_fwd_kernel<true, false, false, false>,  attn_fwd_kernel<true, true, false, false>,
      attn_fwd_kernel<true, false, true, false>,   attn_fwd_kernel<true, true, true, false>,
      attn_fwd_kernel<true, false, false, true>,   attn_fwd_kernel<true, true, false, true>};
  const int vi = scp ? 8 + (drop ? 1 : 0) : (rel ? 4 : 0) + (cz ? 2 : 0) + (drop ? 1 : 0);
  Kern kern = table[vi];
  static bool attr_set[10] = {};
  if (!attr_set[vi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[vi] = true;
  }
  // persistent: one CTA per SM walks a contiguous range of (head, query tile, batch) items
  const int64_t max_ctas = num_sms();
  dim3 grid((unsigned)(n_items < max_ctas ? n_items : max_ctas));
  kern<<<grid, kFThreads, smem_bytes, (cudaStream_t)stream>>>(tq, tk, tv, p);
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
                             const float* scp_table, float* d_scp_table, int64_t scp_q0, int64_t scp_L, int64_t rel_far,
                             void* stream) {
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
  auto al8 = [](int64_t a, int64_t b2, int64_t c) { return a % 8 == 0 && b2 % 8 == 0 && c % 8 == 0; };
  PVQA_REQUIRE(aligned16(o) && aligned16(d_o) && aligned16(dk) && aligned16(dv) && aligned16(dq_accum) &&
                   al8(o_stride_b, o_stride_s, o_stride_h) && al8(do_stride_b, do_stride_s, do_stride_h) &&
                   al8(dk_stride_b, dk_stride_s, dk_stride_h) && al8(dv_stride_b, dv_stride_s, dv_stride_h),
               PVQA_ERR_ALIGN, "attn_bwd: rows must be 16-byte aligned");
  const int n_qt = (int)((Sq + kBM - 1) / kBM);
  const int64_t n_qpad = (int64_t)n_qt * kBM;
  const int64_t n_floats = 2 * kBN + (rel_bias ? 2 * (int64_t)b_rel_copy_stride((int)n_qpad) + (n_qpad + kBN) : 0) + 32 + 512;
  const size_t smem_bytes = 1024 + kBOffFloats + (size_t)n_floats * 4;
  PVQA_REQUIRE(smem_bytes <= 227 * 1024, PVQA_ERR_SHAPE, "attn_bwd: Sq too large for the bias window buffers");
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
  p.lse = lse; p.delta = delta_ws; p.rel_bias = rel_bias; p.key_add = key_add;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.d_rel = d_rel_bias;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.dk_stride_b = dk_stride_b; p.dk_stride_s = dk_stride_s; p.dk_stride_h = dk_stride_h;
  p.dv_stride_b = dv_stride_b; p.dv_stride_s = dv_stride_s; p.dv_stride_h = dv_stride_h;
  p.scale = scale; p.sl2 = scale * kLog2e;
  const uint32_t thr8 = fill_drop(&p.drop, dropout_p, seed, offset, Sk);
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.d_scp = d_scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  p.n_qt = n_qt; p.n_kt = (int)((Sk + kBN - 1) / kBN);
  p.rel_far = (rel_far > 0 && rel_far < (1 << 30)) ? (int)rel_far : 0;
  const int64_t n_items = H * (int64_t)p.n_kt * B;
  PVQA_REQUIRE(n_items < (1ll << 30), PVQA_ERR_SHAPE, "attn_bwd: too many (batch, head, key tile) items");
  p.n_items = (int)n_items;
  const bool rel = rel_bias != nullptr, drop = thr8 != 0, scp = scp_bucket != nullptr, cz = causal != 0;
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                       const AttnBwdParams);
  static const Kern table[10] = {
      attn_bwd_kernel<false, false, false, false>, attn_bwd_kernel<false, true, false, false>,
      attn_bwd_kernel<false, false, true, false>,  attn_bwd_kernel<false, true, true, false>,
      attn_bwd_kernel<true, false, false, false>,  attn_bwd_kernel<true, true, false, false>,
      attn_bwd_kernel<true, false, true, false>,   attn_bwd_kernel<true, true, true, false>,
      attn_bwd_kernel<true, false, false, true>,   attn_bwd_kernel<true, true, false, true>};
  const int vi = scp ? 8 + (drop ? 1 : 0) : (rel ? 4 : 0) + (cz ? 2 : 0) + (drop ? 1 : 0);
  Kern kern = table[vi];
  static bool attr_set[10] = {};
  if (!attr_set[vi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[vi] = true;
  }
  // persistent: one CTA per SM walks a contiguous range of (head, key tile, batch) items.  With waves = k > 1
  // (pvqa_set_attn_bwd_waves, data-parallel runs) the same ranges are cut k times finer and the hardware hands the CTAs
  // to SMs as they free up: an SM that a concurrent NCCL kernel holds for part of the launch then costs its share of
  // the work, not a whole static range queued behind it.
  const int64_t max_ctas = (int64_t)num_sms() * g_attn_bwd_waves.load(std::memory_order_relaxed);
  dim3 grid((unsigned)(n_items < max_ctas ? n_items : max_ctas));
  kern<<<grid, kBThreads, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_bwd");
  return PVQA_OK;
}

