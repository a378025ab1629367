// attn_simt.cu — fp32 attention for the PARITY MODE (compute_dtype = float32).
//
// Same score definition as attn.cu (scale * q.k + rel_bias[j-i] + key_add[j] (+causal), softmax, P@V,
// optional Philox dropout on P), evaluated entirely in fp32 on the CUDA cores so that fp32-mode logits
// stay within 1e-5 of the reference (BASELINE.json north_star); tensor cores would cost TF32/bf16
// rounding.  One warp per query row (forward, dQ) or per key row (dK/dV); no atomics except the
// relative-bias gradient.  This is the correctness path, not the fast path: production training runs
// the tcgen05 kernels in bf16.
#include "common.cuh"

namespace pvqa {

constexpr int kSD = 64;            // head dim
constexpr int kSimtWarps = 4;

struct SimtParams {
  const float* q; const float* k; const float* v;
  float* o; float* lse;
  const float* d_o; const float* delta_in;
  float* delta_out; float* dq; float* dk; float* dv; float* d_rel;
  const float* rel_bias; const float* key_add;
  int B, H, Sq, Sk;
  long long qsb, qss, qsh, ksb, kss, ksh, vsb, vss, vsh, osb, oss, osh;
  long long dosb, doss, dosh, dqsb, dqss, dqsh, dksb, dkss, dksh, dvsb, dvss, dvsh;
  float scale; int causal;
  uint32_t drop_thr8; float drop_scale; uint64_t seed, offset; const unsigned long long* rng_base;
  const uint8_t* scp_bucket; const float* scp_tab; float* d_scp; int scp_q0, scp_L;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
// dot of a 64-float row in global memory with a 64-float row in shared memory
__device__ __forceinline__ float dot64(const float* __restrict__ g, const float* __restrict__ s) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < kSD / 4; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(g) + c);
    const float4 b = reinterpret_cast<const float4*>(s)[c];
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  return acc;
}
__device__ __forceinline__ float score(const SimtParams& p, float dot, int b, int h, int i, int j) {
  float s = dot * p.scale;
  if (p.rel_bias) s += p.rel_bias[(long long)h * (p.Sq + p.Sk - 1) + (j - i + p.Sq - 1)];
  if (p.key_add) s += p.key_add[(long long)b * p.Sk + j];
  if (p.scp_bucket) {
    const int ii = i - p.scp_q0, jj = j - p.scp_q0;
    if (ii >= 0 && ii < p.scp_L && jj >= 0 && jj < p.scp_L)
      s += p.scp_tab[h * 32 + (p.scp_bucket[((long long)b * p.scp_L + ii) * p.scp_L + jj] & 31)];
  }
  if (p.causal && j > i) s = -INFINITY;
  return s;
}
__device__ __forceinline__ float keep_scale(const SimtParams& p, int b, int h, int i, int j) {
  if (!p.drop_thr8) return 1.f;
  const uint64_t grow = ((uint64_t)(b * p.H + h) * p.Sq + i) * (uint64_t)((p.Sk + 15) >> 4);
  const uint32_t keep = attn_dropout_keep16(p.seed, p.offset + (p.rng_base ? *p.rng_base : 0ull), grow + (j >> 4), p.drop_thr8);
  return ((keep >> (j & 15)) & 1u) ? p.drop_scale : 0.f;
}

// ---------------- forward: one warp per query row ----------------
__global__ void __launch_bounds__(kSimtWarps * 32)
attn_f32_fwd_kernel(const SimtParams p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kSimtWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= p.Sq) return;
  float* s_q = sm + warp * (kSD + ((p.Sk + 3) & ~3));      // keep every warp's slice 16-byte aligned
  float* s_p = s_q + kSD;
  const float* qrow = p.q + b * p.qsb + i * p.qss + h * p.qsh;
  s_q[2 * lane] = qrow[2 * lane];
  s_q[2 * lane + 1] = qrow[2 * lane + 1];
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < p.Sk; j += 32) {
    const float s = score(p, dot64(p.k + b * p.ksb + j * p.kss + h * p.ksh, s_q), b, h, i, j);
    s_p[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  const float m_safe = (mx == -INFINITY) ? 0.f : mx;
  float sum = 0.f;
  for (int j = lane; j < p.Sk; j += 32) {
    const float e = expf(s_p[j] - m_safe);
    s_p[j] = e * keep_scale(p, b, h, i, j);
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  const float* vbase = p.v + b * p.vsb + h * p.vsh + 2 * lane;
  for (int j = 0; j < p.Sk; ++j) {
    const float pj = s_p[j];
    const float2 vv = __ldg(reinterpret_cast<const float2*>(vbase + j * p.vss));
    a0 = fmaf(pj, vv.x, a0);
    a1 = fmaf(pj, vv.y, a1);
  }
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  float* orow = p.o + b * p.osb + i * p.oss + h * p.osh;
  orow[2 * lane] = a0 * inv;
  orow[2 * lane + 1] = a1 * inv;
  if (lane == 0 && p.lse) p.lse[((long long)b * p.H + h) * p.Sq + i] = sum > 0.f ? mx + logf(sum) : -INFINITY;
}

// ---------------- backward A: one warp per query row -> dQ, delta, d_rel ----------------
__global__ void __launch_bounds__(kSimtWarps * 32)
attn_f32_bwd_q_kernel(const SimtParams p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kSimtWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= p.Sq) return;
  float* s_q = sm + warp * (2 * kSD + ((p.Sk + 3) & ~3));
  float* s_do = s_q + kSD;
  float* s_ds = s_do + kSD;
  const float* qrow = p.q + b * p.qsb + i * p.qss + h * p.qsh;
  const float* gorow = p.d_o + b * p.dosb + i * p.doss + h * p.dosh;
  const float* orow = p.o + b * p.osb + i * p.oss + h * p.osh;
  const float g0 = gorow[2 * lane], g1 = gorow[2 * lane + 1];
  s_q[2 * lane] = qrow[2 * lane]; s_q[2 * lane + 1] = qrow[2 * lane + 1];
  s_do[2 * lane] = g0; s_do[2 * lane + 1] = g1;
  const float delta = warp_sum(g0 * orow[2 * lane] + g1 * orow[2 * lane + 1]);
  const float lse = p.lse[((long long)b * p.H + h) * p.Sq + i];
  if (lane == 0) p.delta_out[((long long)b * p.H + h) * p.Sq + i] = delta;
  __syncwarp();
  for (int j = lane; j < p.Sk; j += 32) {
    const float s = score(p, dot64(p.k + b * p.ksb + j * p.kss + h * p.ksh, s_q), b, h, i, j);
    const float pr = (lse == -INFINITY || s == -INFINITY) ? 0.f : expf(s - lse);
    const float dP = dot64(p.v + b * p.vsb + j * p.vss + h * p.vsh, s_do) * keep_scale(p, b, h, i, j);
    const float ds = pr * (dP - delta);
    s_ds[j] = ds;
    if (p.d_rel && ds != 0.f) atomicAdd(p.d_rel + (long long)h * (p.Sq + p.Sk - 1) + (j - i + p.Sq - 1), ds);
    if (p.d_scp && ds != 0.f) {
      const int ii = i - p.scp_q0, jj = j - p.scp_q0;
      if (ii >= 0 && ii < p.scp_L && jj >= 0 && jj < p.scp_L)
        atomicAdd(p.d_scp + h * 32 + (p.scp_bucket[((long long)b * p.scp_L + ii) * p.scp_L + jj] & 31), ds);
    }
  }
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  const float* kbase = p.k + b * p.ksb + h * p.ksh + 2 * lane;
  for (int j = 0; j < p.Sk; ++j) {
    const float ds = s_ds[j];
    const float2 kk = __ldg(reinterpret_cast<const float2*>(kbase + j * p.kss));
    a0 = fmaf(ds, kk.x, a0);
    a1 = fmaf(ds, kk.y, a1);
  }
  float* dqrow = p.dq + b * p.dqsb + i * p.dqss + h * p.dqsh;
  dqrow[2 * lane] = a0 * p.scale;
  dqrow[2 * lane + 1] = a1 * p.scale;
}

// ---------------- backward B: one warp per key row -> dK, dV ----------------
__global__ void __launch_bounds__(kSimtWarps * 32)
attn_f32_bwd_kv_kernel(const SimtParams p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * kSimtWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= p.Sk) return;
  float* s_k = sm + warp * (2 * kSD + 2 * ((p.Sq + 3) & ~3));
  float* s_v = s_k + kSD;
  float* s_pd = s_v + kSD;
  float* s_ds = s_pd + ((p.Sq + 3) & ~3);
  const float* krow = p.k + b * p.ksb + j * p.kss + h * p.ksh;
  const float* vrow = p.v + b * p.vsb + j * p.vss + h * p.vsh;
  s_k[2 * lane] = krow[2 * lane]; s_k[2 * lane + 1] = krow[2 * lane + 1];
  s_v[2 * lane] = vrow[2 * lane]; s_v[2 * lane + 1] = vrow[2 * lane + 1];
  __syncwarp();
  for (int i = lane; i < p.Sq; i += 32) {
    const float s = score(p, dot64(p.q + b * p.qsb + i * p.qss + h * p.qsh, s_k), b, h, i, j);
    const float lse = p.lse[((long long)b * p.H + h) * p.Sq + i];
    const float pr = (lse == -INFINITY || s == -INFINITY) ? 0.f : expf(s - lse);
    const float mk = keep_scale(p, b, h, i, j);
    const float dP = dot64(p.d_o + b * p.dosb + i * p.doss + h * p.dosh, s_v) * mk;
    s_pd[i] = pr * mk;
    s_ds[i] = pr * (dP - p.delta_in[((long long)b * p.H + h) * p.Sq + i]);
  }
  __syncwarp();
  float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
  const float* gobase = p.d_o + b * p.dosb + h * p.dosh + 2 * lane;
  const float* qbase = p.q + b * p.qsb + h * p.qsh + 2 * lane;
  for (int i = 0; i < p.Sq; ++i) {
    const float pd = s_pd[i], ds = s_ds[i];
    const float2 g = __ldg(reinterpret_cast<const float2*>(gobase + i * p.doss));
    const float2 qq = __ldg(reinterpret_cast<const float2*>(qbase + i * p.qss));
    v0 = fmaf(pd, g.x, v0); v1 = fmaf(pd, g.y, v1);
    k0 = fmaf(ds, qq.x, k0); k1 = fmaf(ds, qq.y, k1);
  }
  float* dvrow = p.dv + b * p.dvsb + j * p.dvss + h * p.dvsh;
  float* dkrow = p.dk + b * p.dksb + j * p.dkss + h * p.dksh;
  dvrow[2 * lane] = v0; dvrow[2 * lane + 1] = v1;
  dkrow[2 * lane] = k0 * p.scale; dkrow[2 * lane + 1] = k1 * p.scale;
}

static int simt_common_checks(const char* fn, int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t D, int causal,
                              float dropout_p, const int64_t* strides, int n_strides) {
  PVQA_REQUIRE(D == kSD, PVQA_ERR_SHAPE, "%s: head dim %lld unsupported (64 only)", fn, (long long)D);
  PVQA_REQUIRE(B >= 0 && H > 0 && Sq >= 0 && Sk >= 0, PVQA_ERR_SHAPE, "%s: bad dimension", fn);
  PVQA_REQUIRE(!causal || Sq == Sk, PVQA_ERR_SHAPE, "%s: causal requires Sq == Sk", fn);
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "%s: dropout_p must be in [0,1)", fn);
  PVQA_REQUIRE(H <= 65535 && B <= 65535, PVQA_ERR_SHAPE, "%s: H and B must be <= 65535", fn);
  PVQA_REQUIRE(Sq <= 8192 && Sk <= 8192, PVQA_ERR_SHAPE, "%s: sequence too long for the per-warp score buffer", fn);
  for (int s = 0; s < n_strides; ++s)
    PVQA_REQUIRE(strides[s] % 4 == 0, PVQA_ERR_ALIGN, "%s: strides must be multiples of 4 floats (16 bytes)", fn);
  return PVQA_OK;
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_attn_f32_fwd(const float* q, const float* k, const float* v, float* o, float* lse,
                                 const float* rel_bias, const float* key_add, int64_t B, int64_t H, int64_t Sq,
                                 int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                                 int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t v_stride_b,
                                 int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b, int64_t o_stride_s,
                                 int64_t o_stride_h, float scale, int causal, float dropout_p, uint64_t seed,
                                 uint64_t offset, const uint8_t* scp_bucket, const float* scp_table, int64_t scp_q0,
                                 int64_t scp_L, void* stream) {
  const int64_t st[] = {q_stride_b, q_stride_s, q_stride_h, k_stride_b, k_stride_s, k_stride_h,
                        v_stride_b, v_stride_s, v_stride_h, o_stride_b, o_stride_s, o_stride_h};
  int rc = simt_common_checks("attn_f32_fwd", B, H, Sq, Sk, D, causal, dropout_p, st, 12);
  if (rc) return rc;
  if (B == 0 || Sq == 0) return PVQA_OK;
  PVQA_REQUIRE(Sk > 0, PVQA_ERR_SHAPE, "attn_f32_fwd: Sk must be > 0");
  PVQA_REQUIRE(q && k && v && o, PVQA_ERR_NULL, "attn_f32_fwd: NULL pointer");
  PVQA_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), PVQA_ERR_ALIGN,
               "attn_f32_fwd: pointers must be 16-byte aligned");
  SimtParams p{};
  p.q = q; p.k = k; p.v = v; p.o = o; p.lse = lse; p.rel_bias = rel_bias; p.key_add = key_add;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.qsb = q_stride_b; p.qss = q_stride_s; p.qsh = q_stride_h; p.ksb = k_stride_b; p.kss = k_stride_s; p.ksh = k_stride_h;
  p.vsb = v_stride_b; p.vss = v_stride_s; p.vsh = v_stride_h; p.osb = o_stride_b; p.oss = o_stride_s; p.osh = o_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset; p.rng_base = g_rng_base;
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  PVQA_REQUIRE(!scp_bucket || (scp_table && scp_q0 >= 0 && scp_L > 0 && scp_q0 + scp_L <= Sk && Sq == Sk), PVQA_ERR_SHAPE,
               "attn_f32_fwd: bad SCP block");
  const size_t smem = (size_t)kSimtWarps * (kSD + ((Sk + 3) & ~3)) * sizeof(float);
  static size_t smem_cap = 48 * 1024;
  if (smem > smem_cap) {
    cudaError_t e = cudaFuncSetAttribute(attn_f32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_f32_fwd: smem attr: %s", cudaGetErrorString(e));
    smem_cap = smem;
  }
  dim3 grid((unsigned)((Sq + kSimtWarps - 1) / kSimtWarps), (unsigned)H, (unsigned)B);
  attn_f32_fwd_kernel<<<grid, kSimtWarps * 32, smem, (cudaStream_t)stream>>>(p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_f32_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_attn_f32_bwd(const float* q, const float* k, const float* v, const float* o, const float* d_o,
                                 const float* lse, const float* rel_bias, const float* key_add, float* dq, float* dk,
                                 float* dv, float* d_rel_bias, float* delta_ws /* (B,H,Sq) workspace */, int64_t B,
                                 int64_t H, int64_t Sq, int64_t Sk, int64_t D, int64_t q_stride_b, int64_t q_stride_s,
                                 int64_t q_stride_h, int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                                 int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h, int64_t o_stride_b,
                                 int64_t o_stride_s, int64_t o_stride_h, int64_t do_stride_b, int64_t do_stride_s,
                                 int64_t do_stride_h, int64_t dq_stride_b, int64_t dq_stride_s, int64_t dq_stride_h,
                                 int64_t dk_stride_b, int64_t dk_stride_s, int64_t dk_stride_h, int64_t dv_stride_b,
                                 int64_t dv_stride_s, int64_t dv_stride_h, float scale, int causal, float dropout_p,
                                 uint64_t seed, uint64_t offset, const uint8_t* scp_bucket, const float* scp_table,
                                 float* d_scp_table, int64_t scp_q0, int64_t scp_L, void* stream) {
  const int64_t st[] = {q_stride_b, q_stride_s, q_stride_h, k_stride_b, k_stride_s, k_stride_h, v_stride_b,
                        v_stride_s, v_stride_h, o_stride_b, o_stride_s, o_stride_h, do_stride_b, do_stride_s,
                        do_stride_h, dq_stride_b, dq_stride_s, dq_stride_h, dk_stride_b, dk_stride_s, dk_stride_h,
                        dv_stride_b, dv_stride_s, dv_stride_h};
  int rc = simt_common_checks("attn_f32_bwd", B, H, Sq, Sk, D, causal, dropout_p, st, 24);
  if (rc) return rc;
  if (B == 0 || Sq == 0 || Sk == 0) return PVQA_OK;
  PVQA_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && delta_ws, PVQA_ERR_NULL, "attn_f32_bwd: NULL pointer");
  PVQA_REQUIRE(!d_rel_bias || rel_bias, PVQA_ERR_NULL, "attn_f32_bwd: d_rel_bias requested without rel_bias");
  PVQA_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o) && aligned16(dq) &&
                   aligned16(dk) && aligned16(dv),
               PVQA_ERR_ALIGN, "attn_f32_bwd: pointers must be 16-byte aligned");
  SimtParams p{};
  p.q = q; p.k = k; p.v = v; p.o = const_cast<float*>(o); p.lse = const_cast<float*>(lse); p.d_o = d_o;
  p.delta_in = delta_ws; p.delta_out = delta_ws; p.dq = dq; p.dk = dk; p.dv = dv; p.d_rel = d_rel_bias;
  p.rel_bias = rel_bias; p.key_add = key_add;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.qsb = q_stride_b; p.qss = q_stride_s; p.qsh = q_stride_h; p.ksb = k_stride_b; p.kss = k_stride_s; p.ksh = k_stride_h;
  p.vsb = v_stride_b; p.vss = v_stride_s; p.vsh = v_stride_h; p.osb = o_stride_b; p.oss = o_stride_s; p.osh = o_stride_h;
  p.dosb = do_stride_b; p.doss = do_stride_s; p.dosh = do_stride_h; p.dqsb = dq_stride_b; p.dqss = dq_stride_s; p.dqsh = dq_stride_h;
  p.dksb = dk_stride_b; p.dkss = dk_stride_s; p.dksh = dk_stride_h; p.dvsb = dv_stride_b; p.dvss = dv_stride_s; p.dvsh = dv_stride_h;
  p.scale = scale; p.causal = causal;
  p.drop_thr8 = (uint32_t)lrintf(dropout_p * 256.f);
  p.drop_scale = p.drop_thr8 ? 256.f / (256.f - (float)p.drop_thr8) : 1.f;
  p.seed = seed; p.offset = offset; p.rng_base = g_rng_base;
  p.scp_bucket = scp_bucket; p.scp_tab = scp_table; p.d_scp = d_scp_table; p.scp_q0 = (int)scp_q0; p.scp_L = (int)scp_L;
  PVQA_REQUIRE(!scp_bucket || (scp_table && scp_q0 >= 0 && scp_L > 0 && scp_q0 + scp_L <= Sk && Sq == Sk), PVQA_ERR_SHAPE,
               "attn_f32_bwd: bad SCP block");
  cudaStream_t st_ = (cudaStream_t)stream;
  const size_t smem_q = (size_t)kSimtWarps * (2 * kSD + ((Sk + 3) & ~3)) * sizeof(float);
  const size_t smem_kv = (size_t)kSimtWarps * (2 * kSD + 2 * ((Sq + 3) & ~3)) * sizeof(float);
  static size_t cap_q = 48 * 1024, cap_kv = 48 * 1024;
  if (smem_q > cap_q) {
    cudaError_t e = cudaFuncSetAttribute(attn_f32_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_f32_bwd: smem attr: %s", cudaGetErrorString(e));
    cap_q = smem_q;
  }
  if (smem_kv > cap_kv) {
    cudaError_t e = cudaFuncSetAttribute(attn_f32_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "attn_f32_bwd: smem attr: %s", cudaGetErrorString(e));
    cap_kv = smem_kv;
  }
  dim3 gq((unsigned)((Sq + kSimtWarps - 1) / kSimtWarps), (unsigned)H, (unsigned)B);
  attn_f32_bwd_q_kernel<<<gq, kSimtWarps * 32, smem_q, st_>>>(p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_f32_bwd(q)");
  dim3 gk((unsigned)((Sk + kSimtWarps - 1) / kSimtWarps), (unsigned)H, (unsigned)B);
  attn_f32_bwd_kv_kernel<<<gk, kSimtWarps * 32, smem_kv, st_>>>(p);
  count_launch();
  PVQA_CHECK_LAUNCH("attn_f32_bwd(kv)");
  return PVQA_OK;
}
