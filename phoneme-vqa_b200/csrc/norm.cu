// norm.cu — fused normalisation / residual / activation glue of the T5 encoder block and the
// target decoder (HBM-bound element-wise work).
//
// reference semantics restated (no code shared):
//   T5LayerNorm                      transformers modeling_t5.py:46-70   (RMS norm: fp32 variance, no mean, no bias)
//   hidden + dropout(sublayer(...))  transformers modeling_t5.py:T5LayerSelfAttention/T5LayerFF.forward
//   dropout(act(wi(x)))              transformers modeling_t5.py:84-103 (T5DenseActDense), torch TransformerDecoderLayer._ff_block
// Each of these is 4-10 ATen kernels (pow, mean, rsqrt, mul, to, native_dropout, add ...) in the reference;
// here each is one launch reading its operands once.
#include "common.cuh"

namespace pvqa {

constexpr int kNormThreads = 256;
#ifndef PVQA_LN_BWD_CTAS
#define PVQA_LN_BWD_CTAS 2     // resident CTAs per SM of add_dropout_ln_bwd: 128 registers at 2; at 3 (80 registers + 100-350 B
                               // of spills) the encoder launch measured 66 us instead of 60 (gpu call 32) — stays 2
#endif
constexpr int kMaxVec = 8;          // float4 chunks per lane: d <= 32 * 4 * 8 = 1024

__device__ __forceinline__ float warp_sum_n(float x) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float4 r;
  bf16x2_to_f32(u.x, r.x, r.y);
  bf16x2_to_f32(u.y, r.z, r.w);
  return r;
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = f32x2_to_bf16x2(v.x, v.y);
  u.y = f32x2_to_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------- RMS norm forward: one warp per row, the row stays in registers ----------------
template <typename XT, typename YT>
__global__ void __launch_bounds__(kNormThreads)
rms_norm_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ w, YT* __restrict__ y,
                    float* __restrict__ rstd, int N, int d, float eps) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * (kNormThreads / 32) + (threadIdx.x >> 5); row < N; row += gridDim.x * (kNormThreads / 32)) {
    const XT* xr = x + (long long)row * d;
    float4 v[kMaxVec];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
      if ((c * 32 + lane) * 4 < d) {
        v[c] = load4<XT>(xr + (c * 32 + lane) * 4);
        ss += v[c].x * v[c].x + v[c].y * v[c].y + v[c].z * v[c].z + v[c].w * v[c].w;
      }
    }
    ss = warp_sum_n(ss);
    const float r = rsqrtf(ss / (float)d + eps);
    if (lane == 0 && rstd) rstd[row] = r;
    YT* yr = y + (long long)row * d;
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
      if ((c * 32 + lane) * 4 < d) {
        const float4 ww = *reinterpret_cast<const float4*>(w + (c * 32 + lane) * 4);
        float4 o;
        o.x = ww.x * (v[c].x * r); o.y = ww.y * (v[c].y * r); o.z = ww.z * (v[c].z * r); o.w = ww.w * (v[c].w * r);
        store4<YT>(yr + (c * 32 + lane) * 4, o);
      }
    }
  }
}

// ---------------- RMS norm backward ----------------
// dx = rstd * (g - xhat * mean(g * xhat)),  g = dy * w,  xhat = x * rstd;   dw += sum_rows dy * xhat
template <typename XT, typename YT>
__global__ void __launch_bounds__(kNormThreads)
rms_norm_bwd_kernel(const YT* __restrict__ dy, const XT* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ rstd, const XT* __restrict__ d_res, XT* __restrict__ dx,
                    float* __restrict__ dw, int N, int d) {
  extern __shared__ float s_dw[];             // [d] block partial
  const int lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < d; c += kNormThreads) s_dw[c] = 0.f;
  __syncthreads();
  float4 acc[kMaxVec];
#pragma unroll
  for (int c = 0; c < kMaxVec; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * (kNormThreads / 32) + (threadIdx.x >> 5); row < N; row += gridDim.x * (kNormThreads / 32)) {
    const float r = rstd[row];
    const XT* xr = x + (long long)row * d;
    const YT* gr = dy + (long long)row * d;
    float4 xh[kMaxVec], g[kMaxVec];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
      if ((c * 32 + lane) * 4 < d) {
        const int col = (c * 32 + lane) * 4;
        const float4 xv = load4<XT>(xr + col);
        const float4 gy = load4<YT>(gr + col);
        const float4 ww = *reinterpret_cast<const float4*>(w + col);
        xh[c] = make_float4(xv.x * r, xv.y * r, xv.z * r, xv.w * r);
        g[c] = make_float4(gy.x * ww.x, gy.y * ww.y, gy.z * ww.z, gy.w * ww.w);
        dot += g[c].x * xh[c].x + g[c].y * xh[c].y + g[c].z * xh[c].z + g[c].w * xh[c].w;
        acc[c].x += gy.x * xh[c].x; acc[c].y += gy.y * xh[c].y; acc[c].z += gy.z * xh[c].z; acc[c].w += gy.w * xh[c].w;
      }
    }
    dot = warp_sum_n(dot) / (float)d;
    XT* dr = dx + (long long)row * d;
    const XT* rr = d_res ? d_res + (long long)row * d : nullptr;
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
      if ((c * 32 + lane) * 4 < d) {
        float4 o;
        o.x = r * (g[c].x - xh[c].x * dot); o.y = r * (g[c].y - xh[c].y * dot);
        o.z = r * (g[c].z - xh[c].z * dot); o.w = r * (g[c].w - xh[c].w * dot);
        if (rr) {          // fused residual-gradient accumulation: dx = d_residual + d(norm branch)
          const float4 e = load4<XT>(rr + (c * 32 + lane) * 4);
          o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
        }
        store4<XT>(dr + (c * 32 + lane) * 4, o);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxVec; ++c) {
    if ((c * 32 + lane) * 4 < d) {
      const int col = (c * 32 + lane) * 4;
      atomicAdd(s_dw + col, acc[c].x); atomicAdd(s_dw + col + 1, acc[c].y);
      atomicAdd(s_dw + col + 2, acc[c].z); atomicAdd(s_dw + col + 3, acc[c].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kNormThreads) atomicAdd(dw + c, s_dw[c]);
}

// ---------------- out = hidden + dropout(update) ----------------
template <typename UT>
__global__ void __launch_bounds__(kNormThreads)
residual_dropout_add_kernel(const float* __restrict__ hidden, const UT* __restrict__ upd, float* __restrict__ out,
                            long long n8, uint32_t thr16, float scale, uint64_t seed, uint64_t offset, const unsigned long long* rng_base) {
  if (thr16 && rng_base) offset += *rng_base;
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kNormThreads) {
    f8 h = Vec8<float>::load_stream(hidden + i * 8);
    f8 u = Vec8<UT>::load_stream(upd + i * 8);
    if (thr16) {
      const uint32_t m = dropout_keep8(seed, offset, (uint64_t)i * 8, thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) h.v[j] += ((m >> j) & 1u) ? u.v[j] * scale : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) h.v[j] += u.v[j];
    }
    Vec8<float>::store(out + i * 8, h);
  }
}
// d_update = d_out * mask * scale (cast to the update's dtype); d_hidden is d_out itself.
template <typename UT>
__global__ void __launch_bounds__(kNormThreads)
residual_dropout_bwd_kernel(const float* __restrict__ d_out, UT* __restrict__ d_upd, long long n8, uint32_t thr16,
                            float scale, uint64_t seed, uint64_t offset, const unsigned long long* rng_base) {
  if (thr16 && rng_base) offset += *rng_base;
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kNormThreads) {
    f8 g = Vec8<float>::load_stream(d_out + i * 8);
    if (thr16) {
      const uint32_t m = dropout_keep8(seed, offset, (uint64_t)i * 8, thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] = ((m >> j) & 1u) ? g.v[j] * scale : 0.f;
    }
    Vec8<UT>::store(d_upd + i * 8, g);
  }
}

// ---------------- y = dropout(relu(x)) in place-able form; backward needs only y ----------------
// The kernel was issue-bound on its random numbers (ncu: 53 % issue-active, 36 % DRAM with one Philox4x32-10 call and
// eight 16-bit compares per 8 elements), so a thread now owns 16 consecutive elements and spends ONE Philox4x32-7 call
// on them: 8 random bits per element, keep iff byte >= thr8 (the drop probability is thr8 / 256, p quantised to 1/256
// like the attention kernels; the 1/keep scale uses the quantised value, and the backward takes it from the same
// drop_consts8).  The mask is never needed again: the backward reads it off y != 0.
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
relu_dropout_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8, uint32_t thr8, float scale,
                        uint64_t seed, uint64_t offset, const unsigned long long* rng_base) {
  if (thr8 && rng_base) offset += *rng_base;
  const long long n16 = (n8 + 1) >> 1;                     // units of 16 elements; the last one may hold only 8
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < n16; i += (long long)gridDim.x * kNormThreads) {
    const bool two = 2 * i + 1 < n8;
    f8 a = Vec8<T>::load_stream(x + i * 16);
    f8 b{};
    if (two) b = Vec8<T>::load_stream(x + i * 16 + 8);
    uint32_t w[4] = {~0u, ~0u, ~0u, ~0u};
    if (thr8) {
      const uint64_t c = (uint64_t)i + offset;
      const uint4 r = philox4x32_r<7>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5A19u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ka = ((w[j >> 2] >> (8 * (j & 3))) & 0xffu) >= thr8;
      const bool kb = ((w[2 + (j >> 2)] >> (8 * (j & 3))) & 0xffu) >= thr8;
      a.v[j] = (ka && a.v[j] > 0.f) ? a.v[j] * scale : 0.f;
      b.v[j] = (kb && b.v[j] > 0.f) ? b.v[j] * scale : 0.f;
    }
    Vec8<T>::store(y + i * 16, a);
    if (two) Vec8<T>::store(y + i * 16 + 8, b);
  }
}
// dx = (y != 0) ? dy * scale : 0     (y != 0 <=> x > 0 and kept)
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
relu_dropout_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n8, float scale) {
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kNormThreads) {
    f8 g = Vec8<T>::load_stream(dy + i * 8);
    f8 v = Vec8<T>::load_stream(y + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] = (v.v[j] != 0.f) ? g.v[j] * scale : 0.f;
    Vec8<T>::store(dx + i * 8, g);
  }
}

// ---------------- y = LayerNorm(hidden + dropout(update)): post-norm tail of the target decoder layer ----------------
// One warp per row, the row lives in registers (NV chunks of 8 elements per lane: d <= 256 * NV).  The dropout
// mask uses the same flat element index (row * d + col) as residual_dropout_add, so the fused kernel is
// bit-compatible with the unfused pair.  Outputs: z (pre-norm sum, saved for backward), y (fp32 residual
// stream) and optionally y_lp (the low-precision copy the next GEMM reads).
// 4 CTAs per SM (64 registers, no spills up to NV = 3): the kernel is latency-bound on its loads (ncu: long
// scoreboard, 31 % warps active at 3 CTAs per SM), and 8 CTAs per SM of grid then make exactly two waves.
template <typename HT, typename UT, typename LT, int NV>
__global__ void __launch_bounds__(kNormThreads, 4)
add_dropout_ln_fwd_kernel(const HT* __restrict__ hidden, const UT* __restrict__ upd, const float* __restrict__ gamma,
                          const float* __restrict__ beta, HT* __restrict__ z, float* __restrict__ y,
                          LT* __restrict__ y_lp, float* __restrict__ mean, float* __restrict__ rstd, int N, int d,
                          float eps, uint32_t thr16, float scale, uint64_t seed, uint64_t offset,
                          const unsigned long long* rng_base, int rms) {
  if (thr16 && rng_base) offset += *rng_base;
  const int lane = threadIdx.x & 31;
  const float inv_d = 1.0f / (float)d;
  for (int row = blockIdx.x * (kNormThreads / 32) + (threadIdx.x >> 5); row < N; row += gridDim.x * (kNormThreads / 32)) {
    const long long base = (long long)row * d;
    f8 v[NV];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        v[c] = Vec8<HT>::load_stream(hidden + base + col);
        if (upd) {
          const f8 u = Vec8<UT>::load_stream(upd + base + col);
          if (thr16) {
            const uint32_t m = dropout_keep8(seed, offset, (uint64_t)(base + col), thr16);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[c].v[j] += ((m >> j) & 1u) ? u.v[j] * scale : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[c].v[j] += u.v[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[c].v[j];
      }
    }
    const float mu = rms ? 0.f : warp_sum_n(s) * inv_d;      // rms: T5LayerNorm (no mean, no bias)
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if ((c * 32 + lane) * 8 < d) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[c].v[j] - mu; ss += t * t; }
      }
    }
    const float r = rsqrtf(warp_sum_n(ss) * inv_d + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = r;
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        if (z) Vec8<HT>::store(z + base + col, v[c]);
        const f8 g = Vec8<float>::load(gamma + col);
        f8 o;
        if (beta) {
          const f8 bt = Vec8<float>::load(beta + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = (v[c].v[j] - mu) * r * g.v[j] + bt.v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = g.v[j] * ((v[c].v[j] - mu) * r);
        }
        if (y) Vec8<float>::store(y + base + col, o);
        if (y_lp) Vec8<LT>::store(y_lp + base + col, o);
      }
    }
  }
}

// backward of the fused tail.  g = dy (+ dy_lp), xhat = (z - mean) * rstd, wg = g * gamma:
//   dz = rstd * (wg - mean(wg) - xhat * mean(wg * xhat));  d_hidden = dz;  d_update = mask * dz / keep
//   dgamma += sum_rows g * xhat;  dbeta += sum_rows g      (register partials -> smem -> one global atomic per CTA column)
template <typename UT, typename LT, int NV, bool LN>      // LN: LayerNorm (mean, beta); otherwise T5 RMS norm
__global__ void __launch_bounds__(kNormThreads, PVQA_LN_BWD_CTAS)
add_dropout_ln_bwd_kernel(const float* __restrict__ dy, const LT* __restrict__ dy_lp, const float* __restrict__ d_res,
                          const float* __restrict__ z,
                          const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                          float* __restrict__ d_hidden, UT* __restrict__ d_upd, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, int N, int d, uint32_t thr16, float scale, uint64_t seed,
                          uint64_t offset, const unsigned long long* rng_base) {
  extern __shared__ float s_part[];           // [2][d] block partials of dgamma, dbeta
  if (thr16 && rng_base) offset += *rng_base;
  const int lane = threadIdx.x & 31;
  const float inv_d = 1.0f / (float)d;
  for (int c = threadIdx.x; c < 2 * d; c += kNormThreads) s_part[c] = 0.f;
  __syncthreads();
  f8 acc_g[NV], acc_b[LN ? NV : 1];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc_g[c].v[j] = 0.f;
      if (LN) acc_b[c].v[j] = 0.f;
    }
  }
  const int row_stride = gridDim.x * (kNormThreads / 32);
  for (int row = blockIdx.x * (kNormThreads / 32) + (threadIdx.x >> 5); row < N; row += row_stride) {
    const long long base = (long long)row * d;
    // The kernel keeps its column sums in registers, so few warps walk many rows (2 CTAs per SM, ~9 rows per warp) and
    // every row used to pay two dependent DRAM round trips (z / dy, then d_res after the warp reduction).  The next
    // row's operands are pulled into L2 while this row is computed: prefetches need no destination registers.
    if (row + row_stride < N) {
      const long long nb = (long long)(row + row_stride) * d;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < d) {
          prefetch_l2(z + nb + col);
          if (dy) prefetch_l2(dy + nb + col);
          if (dy_lp) prefetch_l2(dy_lp + nb + col);
          if (d_res) prefetch_l2(d_res + nb + col);
        }
      }
    }
    const float mu = LN ? mean[row] : 0.f, r = rstd[row];
    f8 xh[NV], wg[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        f8 g;
        if (dy) g = Vec8<float>::load_stream(dy + base + col);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) g.v[j] = 0.f;
        }
        if (dy_lp) {
          const f8 g2 = Vec8<LT>::load_stream(dy_lp + base + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
        }
        const f8 zz = Vec8<float>::load_stream(z + base + col);
        const f8 gm = Vec8<float>::load(gamma + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (zz.v[j] - mu) * r;
          xh[c].v[j] = xhat;
          acc_g[c].v[j] = fmaf(g.v[j], xhat, acc_g[c].v[j]);
          if (LN) acc_b[c].v[j] += g.v[j];
          const float w = g.v[j] * gm.v[j];
          wg[c].v[j] = w;
          if (LN) s1 += w;
          s2 = fmaf(w, xhat, s2);
        }
      }
    }
    const float c1 = LN ? warp_sum_n(s1) * inv_d : 0.f, c2 = warp_sum_n(s2) * inv_d;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        f8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = r * (wg[c].v[j] - c1 - xh[c].v[j] * c2);
        if (d_res) {         // pre-norm residual path: the gradient that bypasses the norm
          const f8 e = Vec8<float>::load_stream(d_res + base + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] += e.v[j];
        }
        Vec8<float>::store(d_hidden + base + col, o);
        if (d_upd) {
          if (thr16) {
            const uint32_t m = dropout_keep8(seed, offset, (uint64_t)(base + col), thr16);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = ((m >> j) & 1u) ? o.v[j] * scale : 0.f;
          }
          Vec8<UT>::store(d_upd + base + col, o);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(s_part + col + j, acc_g[c].v[j]);
        if (LN) atomicAdd(s_part + d + col + j, acc_b[c].v[j]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kNormThreads) {
    atomicAdd(dgamma + c, s_part[c]);
    if (LN) atomicAdd(dbeta + c, s_part[d + c]);
  }
}

// ---------------- out[c] += sum_rows x[r][c]: bias gradient of a Linear (fp32 accumulation) ----------------
// CTA = 32 column groups (8 columns each) x 8 row lanes; grid = (column slabs, row slabs).
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
col_sum_kernel(const T* __restrict__ x, float* __restrict__ out, int N, int d, int rows_per_cta) {
  __shared__ float s_acc[8][32 * 8 + 1];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  f8 acc;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
  if (col < d) {
    int r = r0 + rl;
    for (; r + 24 < r1; r += 32) {          // four independent loads in flight per thread
      const f8 a = Vec8<T>::load_stream(x + (long long)r * d + col);
      const f8 b = Vec8<T>::load_stream(x + (long long)(r + 8) * d + col);
      const f8 c = Vec8<T>::load_stream(x + (long long)(r + 16) * d + col);
      const f8 e = Vec8<T>::load_stream(x + (long long)(r + 24) * d + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] += (a.v[j] + b.v[j]) + (c.v[j] + e.v[j]);
    }
    for (; r < r1; r += 8) {
      const f8 a = Vec8<T>::load_stream(x + (long long)r * d + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] += a.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_acc[rl][cg * 8 + j] = acc.v[j];
  __syncthreads();
  {
    const int c = threadIdx.x;              // 256 threads == 256 columns of this slab
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_acc[k][c];
    const int gc = blockIdx.x * 256 + c;
    if (gc < d) atomicAdd(out + gc, t);
  }
}

// ---------------- dst[r][0:d] (row stride ld elements, T) = src[r][0:d] (fp32 contiguous) ----------------
// the fp32 dQ accumulator of the attention backward cast into the q slot of the packed (B,S,3,H,D) gradient
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
cast_rows_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n8, int d8, long long ld) {
  for (long long i = (long long)blockIdx.x * kNormThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kNormThreads) {
    const long long r = i / d8;
    const int c = (int)(i - r * d8) * 8;
    Vec8<T>::store(dst + r * ld + c, Vec8<float>::load_stream(src + i * 8));
  }
}

static int ew_grid(long long n8) {
  long long need = (n8 + kNormThreads - 1) / kNormThreads;
  long long cap = (long long)num_sms() * 8;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}
static void drop_consts(float p, uint32_t& thr16, float& scale) {
  thr16 = p > 0.f ? (uint32_t)lrintf(p * 65536.f) : 0u;
  scale = thr16 ? 65536.f / (65536.f - (float)thr16) : 1.f;
}
// 8-bit decisions (relu_dropout): p quantised to 1/256, never rounded down to "no dropout" for p > 0
static void drop_consts8(float p, uint32_t& thr8, float& scale) {
  thr8 = p > 0.f ? (uint32_t)lrintf(p * 256.f) : 0u;
  if (p > 0.f && thr8 == 0) thr8 = 1;
  if (thr8 > 255) thr8 = 255;
  scale = thr8 ? 256.f / (256.f - (float)thr8) : 1.f;
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_rms_norm_fwd(const void* x, const float* w, void* y, float* rstd, int64_t N, int64_t d,
                                 float eps, int x_dtype, int y_dtype, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "rms_norm_fwd: bad dimension");
  PVQA_REQUIRE(d % 4 == 0 && d <= 128 * kMaxVec, PVQA_ERR_SHAPE,
               "rms_norm_fwd: d=%lld must be a multiple of 4 and <= %d", (long long)d, 128 * kMaxVec);
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(x && w && y, PVQA_ERR_NULL, "rms_norm_fwd: NULL pointer");
  PVQA_REQUIRE(aligned32(x) && aligned32(w) && aligned32(y), PVQA_ERR_ALIGN, "rms_norm_fwd: 32-byte alignment required");
  const int warps = kNormThreads / 32;
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == PVQA_F32 && y_dtype == PVQA_BF16)
    rms_norm_fwd_kernel<float, __nv_bfloat16><<<grid, kNormThreads, 0, st>>>((const float*)x, w, (__nv_bfloat16*)y, rstd, (int)N, (int)d, eps);
  else if (x_dtype == PVQA_F32 && y_dtype == PVQA_F32)
    rms_norm_fwd_kernel<float, float><<<grid, kNormThreads, 0, st>>>((const float*)x, w, (float*)y, rstd, (int)N, (int)d, eps);
  else if (x_dtype == PVQA_BF16 && y_dtype == PVQA_BF16)
    rms_norm_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kNormThreads, 0, st>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, rstd, (int)N, (int)d, eps);
  else
    return fail(PVQA_ERR_DTYPE, "rms_norm_fwd: unsupported dtype combination");
  count_launch();
  PVQA_CHECK_LAUNCH("rms_norm_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_rms_norm_bwd(const void* dy, const void* x, const float* w, const float* rstd,
                                 const void* d_residual /* optional, x dtype: added to dx */, void* dx,
                                 float* dw /* accumulated */, int64_t N, int64_t d, int x_dtype, int y_dtype,
                                 void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "rms_norm_bwd: bad dimension");
  PVQA_REQUIRE(d % 4 == 0 && d <= 128 * kMaxVec, PVQA_ERR_SHAPE, "rms_norm_bwd: d must be a multiple of 4 and <= %d", 128 * kMaxVec);
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(dy && x && w && rstd && dx && dw, PVQA_ERR_NULL, "rms_norm_bwd: NULL pointer");
  PVQA_REQUIRE(aligned32(dy) && aligned32(x) && aligned32(w) && aligned32(dx), PVQA_ERR_ALIGN, "rms_norm_bwd: 32-byte alignment required");
  const int warps = kNormThreads / 32;
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * 2;   // 128 regs -> 2 CTAs/SM resident
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)d * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == PVQA_F32 && y_dtype == PVQA_BF16)
    rms_norm_bwd_kernel<float, __nv_bfloat16><<<grid, kNormThreads, smem, st>>>((const __nv_bfloat16*)dy, (const float*)x, w, rstd, (const float*)d_residual, (float*)dx, dw, (int)N, (int)d);
  else if (x_dtype == PVQA_F32 && y_dtype == PVQA_F32)
    rms_norm_bwd_kernel<float, float><<<grid, kNormThreads, smem, st>>>((const float*)dy, (const float*)x, w, rstd, (const float*)d_residual, (float*)dx, dw, (int)N, (int)d);
  else if (x_dtype == PVQA_BF16 && y_dtype == PVQA_BF16)
    rms_norm_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kNormThreads, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, w, rstd, (const __nv_bfloat16*)d_residual, (__nv_bfloat16*)dx, dw, (int)N, (int)d);
  else
    return fail(PVQA_ERR_DTYPE, "rms_norm_bwd: unsupported dtype combination");
  count_launch();
  PVQA_CHECK_LAUNCH("rms_norm_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_residual_dropout_add(const float* hidden, const void* update, float* out, int64_t n,
                                         int upd_dtype, float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(n >= 0 && n % 8 == 0, PVQA_ERR_SHAPE, "residual_dropout_add: element count must be a multiple of 8");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "residual_dropout_add: dropout_p must be in [0,1)");
  if (n == 0) return PVQA_OK;
  PVQA_REQUIRE(hidden && update && out, PVQA_ERR_NULL, "residual_dropout_add: NULL pointer");
  PVQA_REQUIRE(aligned32(hidden) && aligned32(update) && aligned32(out), PVQA_ERR_ALIGN, "residual_dropout_add: 32-byte alignment required");
  uint32_t thr; float sc;
  drop_consts(dropout_p, thr, sc);
  cudaStream_t st = (cudaStream_t)stream;
  if (upd_dtype == PVQA_BF16)
    residual_dropout_add_kernel<__nv_bfloat16><<<ew_grid(n / 8), kNormThreads, 0, st>>>(hidden, (const __nv_bfloat16*)update, out, n / 8, thr, sc, seed, offset, g_rng_base);
  else if (upd_dtype == PVQA_F32)
    residual_dropout_add_kernel<float><<<ew_grid(n / 8), kNormThreads, 0, st>>>(hidden, (const float*)update, out, n / 8, thr, sc, seed, offset, g_rng_base);
  else
    return fail(PVQA_ERR_DTYPE, "residual_dropout_add: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("residual_dropout_add");
  return PVQA_OK;
}

extern "C" int pvqa_residual_dropout_bwd(const float* d_out, void* d_update, int64_t n, int upd_dtype, float dropout_p,
                                         uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(n >= 0 && n % 8 == 0, PVQA_ERR_SHAPE, "residual_dropout_bwd: element count must be a multiple of 8");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "residual_dropout_bwd: dropout_p must be in [0,1)");
  if (n == 0) return PVQA_OK;
  PVQA_REQUIRE(d_out && d_update, PVQA_ERR_NULL, "residual_dropout_bwd: NULL pointer");
  PVQA_REQUIRE(aligned32(d_out) && aligned32(d_update), PVQA_ERR_ALIGN, "residual_dropout_bwd: 32-byte alignment required");
  uint32_t thr; float sc;
  drop_consts(dropout_p, thr, sc);
  cudaStream_t st = (cudaStream_t)stream;
  if (upd_dtype == PVQA_BF16)
    residual_dropout_bwd_kernel<__nv_bfloat16><<<ew_grid(n / 8), kNormThreads, 0, st>>>(d_out, (__nv_bfloat16*)d_update, n / 8, thr, sc, seed, offset, g_rng_base);
  else if (upd_dtype == PVQA_F32)
    residual_dropout_bwd_kernel<float><<<ew_grid(n / 8), kNormThreads, 0, st>>>(d_out, (float*)d_update, n / 8, thr, sc, seed, offset, g_rng_base);
  else
    return fail(PVQA_ERR_DTYPE, "residual_dropout_bwd: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("residual_dropout_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_relu_dropout_fwd(const void* x, void* y, int64_t n, int dtype, float dropout_p, uint64_t seed,
                                     uint64_t offset, void* stream) {
  PVQA_REQUIRE(n >= 0 && n % 8 == 0, PVQA_ERR_SHAPE, "relu_dropout_fwd: element count must be a multiple of 8");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "relu_dropout_fwd: dropout_p must be in [0,1)");
  if (n == 0) return PVQA_OK;
  PVQA_REQUIRE(x && y, PVQA_ERR_NULL, "relu_dropout_fwd: NULL pointer");
  PVQA_REQUIRE(aligned32(x) && aligned32(y), PVQA_ERR_ALIGN, "relu_dropout_fwd: 32-byte alignment required");
  uint32_t thr; float sc;
  drop_consts8(dropout_p, thr, sc);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid((n / 8 + 1) / 2);
  if (dtype == PVQA_BF16)
    relu_dropout_fwd_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8, thr, sc, seed, offset, g_rng_base);
  else if (dtype == PVQA_F32)
    relu_dropout_fwd_kernel<float><<<grid, kNormThreads, 0, st>>>((const float*)x, (float*)y, n / 8, thr, sc, seed, offset, g_rng_base);
  else
    return fail(PVQA_ERR_DTYPE, "relu_dropout_fwd: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("relu_dropout_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_relu_dropout_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, float dropout_p,
                                     void* stream) {
  PVQA_REQUIRE(n >= 0 && n % 8 == 0, PVQA_ERR_SHAPE, "relu_dropout_bwd: element count must be a multiple of 8");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "relu_dropout_bwd: dropout_p must be in [0,1)");
  if (n == 0) return PVQA_OK;
  PVQA_REQUIRE(dy && y && dx, PVQA_ERR_NULL, "relu_dropout_bwd: NULL pointer");
  PVQA_REQUIRE(aligned32(dy) && aligned32(y) && aligned32(dx), PVQA_ERR_ALIGN, "relu_dropout_bwd: 32-byte alignment required");
  uint32_t thr; float sc;
  drop_consts8(dropout_p, thr, sc);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PVQA_BF16)
    relu_dropout_bwd_kernel<__nv_bfloat16><<<ew_grid(n / 8), kNormThreads, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, n / 8, sc);
  else if (dtype == PVQA_F32)
    relu_dropout_bwd_kernel<float><<<ew_grid(n / 8), kNormThreads, 0, st>>>((const float*)dy, (const float*)y, (float*)dx, n / 8, sc);
  else
    return fail(PVQA_ERR_DTYPE, "relu_dropout_bwd: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("relu_dropout_bwd");
  return PVQA_OK;
}

// ---- fused post-norm tail: y = LayerNorm(hidden + dropout(update)) ----
template <typename HT, typename UT, typename LT>
static int launch_add_ln_fwd(int nv, int grid, cudaStream_t st, const HT* hidden, const UT* upd, const float* gamma,
                             const float* beta, HT* z, float* y, LT* y_lp, float* mean, float* rstd, int N, int d,
                             float eps, uint32_t thr, float sc, uint64_t seed, uint64_t offset, int rms = 0) {
#define PVQA_LN_FWD(NV) add_dropout_ln_fwd_kernel<HT, UT, LT, NV><<<grid, kNormThreads, 0, st>>>( \
    hidden, upd, gamma, beta, z, y, y_lp, mean, rstd, N, d, eps, thr, sc, seed, offset, g_rng_base, rms)
  switch (nv) { case 1: PVQA_LN_FWD(1); break; case 2: PVQA_LN_FWD(2); break; case 3: PVQA_LN_FWD(3); break; default: PVQA_LN_FWD(4); }
#undef PVQA_LN_FWD
  return PVQA_OK;
}
template <typename UT, typename LT>
static int launch_add_ln_bwd(int nv, int grid, size_t smem, cudaStream_t st, const float* dy, const LT* dy_lp,
                             const float* d_res, const float* z, const float* gamma, const float* mean, const float* rstd, float* d_hidden,
                             UT* d_upd, float* dgamma, float* dbeta, int N, int d, uint32_t thr, float sc, uint64_t seed,
                             uint64_t offset) {
#define PVQA_LN_BWD(NV, LN) add_dropout_ln_bwd_kernel<UT, LT, NV, LN><<<grid, kNormThreads, smem, st>>>( \
    dy, dy_lp, d_res, z, gamma, mean, rstd, d_hidden, d_upd, dgamma, dbeta, N, d, thr, sc, seed, offset, g_rng_base)
  if (mean != nullptr && dbeta != nullptr) {
    switch (nv) { case 1: PVQA_LN_BWD(1, true); break; case 2: PVQA_LN_BWD(2, true); break; case 3: PVQA_LN_BWD(3, true); break; default: PVQA_LN_BWD(4, true); }
  } else {
    switch (nv) { case 1: PVQA_LN_BWD(1, false); break; case 2: PVQA_LN_BWD(2, false); break; case 3: PVQA_LN_BWD(3, false); break; default: PVQA_LN_BWD(4, false); }
  }
#undef PVQA_LN_BWD
  return PVQA_OK;
}

extern "C" int pvqa_add_dropout_ln_fwd(const void* hidden, int hidden_dtype, const void* update, int upd_dtype,
                                       const float* gamma, const float* beta, void* z, float* y, void* y_lp, int lp_dtype, float* mean,
                                       float* rstd, int64_t N, int64_t d, float eps, float dropout_p, uint64_t seed,
                                       uint64_t offset, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "add_dropout_ln_fwd: bad dimension");
  PVQA_REQUIRE(d % 8 == 0 && d <= 1024, PVQA_ERR_SHAPE, "add_dropout_ln_fwd: d=%lld must be a multiple of 8 and <= 1024", (long long)d);
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "add_dropout_ln_fwd: dropout_p must be in [0,1)");
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(hidden && gamma && beta && (y || y_lp), PVQA_ERR_NULL, "add_dropout_ln_fwd: NULL pointer");  // beta required: LayerNorm
  PVQA_REQUIRE(aligned32(hidden) && aligned32(update) && aligned32(gamma) && aligned32(beta) && aligned32(z) &&
                   aligned32(y) && aligned32(y_lp), PVQA_ERR_ALIGN, "add_dropout_ln_fwd: 32-byte alignment required");
  PVQA_REQUIRE(!y_lp || lp_dtype == PVQA_BF16, PVQA_ERR_DTYPE, "add_dropout_ln_fwd: the low-precision copy must be bf16");
  uint32_t thr; float sc;
  drop_consts(update ? dropout_p : 0.f, thr, sc);
  const int warps = kNormThreads / 32, nv = (int)((d + 255) / 256);
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  if (hidden_dtype == PVQA_F32 && upd_dtype == PVQA_BF16)
    launch_add_ln_fwd<float, bf, bf>(nv, grid, st, (const float*)hidden, (const bf*)update, gamma, beta, (float*)z, y, (bf*)y_lp, mean, rstd, (int)N, (int)d, eps, thr, sc, seed, offset);
  else if (hidden_dtype == PVQA_F32 && upd_dtype == PVQA_F32)
    launch_add_ln_fwd<float, float, bf>(nv, grid, st, (const float*)hidden, (const float*)update, gamma, beta, (float*)z, y, (bf*)y_lp, mean, rstd, (int)N, (int)d, eps, thr, sc, seed, offset);
  else if (hidden_dtype == PVQA_BF16 && upd_dtype == PVQA_BF16)     // frozen bf16 ViT tower (pre-norm, no dropout)
    launch_add_ln_fwd<bf, bf, bf>(nv, grid, st, (const bf*)hidden, (const bf*)update, gamma, beta, (bf*)z, y, (bf*)y_lp, mean, rstd, (int)N, (int)d, eps, thr, sc, seed, offset);
  else
    return fail(PVQA_ERR_DTYPE, "add_dropout_ln_fwd: unsupported hidden/update dtype combination");
  count_launch();
  PVQA_CHECK_LAUNCH("add_dropout_ln_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_add_dropout_ln_bwd(const float* dy, const void* dy_lp, int lp_dtype, const float* z,
                                       const float* gamma, const float* mean, const float* rstd, float* d_hidden,
                                       void* d_update, int upd_dtype, float* dgamma, float* dbeta, int64_t N, int64_t d,
                                       float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "add_dropout_ln_bwd: bad dimension");
  PVQA_REQUIRE(d % 8 == 0 && d <= 1024, PVQA_ERR_SHAPE, "add_dropout_ln_bwd: d must be a multiple of 8 and <= 1024");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "add_dropout_ln_bwd: dropout_p must be in [0,1)");
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE((dy || dy_lp) && z && gamma && mean && rstd && d_hidden && dgamma && dbeta, PVQA_ERR_NULL, "add_dropout_ln_bwd: NULL pointer");
  PVQA_REQUIRE(aligned32(dy) && aligned32(dy_lp) && aligned32(z) && aligned32(gamma) && aligned32(d_hidden) && aligned32(d_update),
               PVQA_ERR_ALIGN, "add_dropout_ln_bwd: 16-byte alignment required");
  PVQA_REQUIRE(!dy_lp || lp_dtype == PVQA_BF16, PVQA_ERR_DTYPE, "add_dropout_ln_bwd: the low-precision gradient must be bf16");
  uint32_t thr; float sc;
  drop_consts(dropout_p, thr, sc);
  const int warps = kNormThreads / 32, nv = (int)((d + 255) / 256);
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * PVQA_LN_BWD_CTAS;
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)2 * d * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (upd_dtype == PVQA_BF16)
    launch_add_ln_bwd<__nv_bfloat16, __nv_bfloat16>(nv, grid, smem, st, dy, (const __nv_bfloat16*)dy_lp, nullptr, z, gamma, mean, rstd, d_hidden, (__nv_bfloat16*)d_update, dgamma, dbeta, (int)N, (int)d, thr, sc, seed, offset);
  else if (upd_dtype == PVQA_F32)
    launch_add_ln_bwd<float, __nv_bfloat16>(nv, grid, smem, st, dy, (const __nv_bfloat16*)dy_lp, nullptr, z, gamma, mean, rstd, d_hidden, (float*)d_update, dgamma, dbeta, (int)N, (int)d, thr, sc, seed, offset);
  else
    return fail(PVQA_ERR_DTYPE, "add_dropout_ln_bwd: bad update dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("add_dropout_ln_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_col_sum(const void* x, float* out, int64_t N, int64_t d, int dtype, int accumulate, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "col_sum: bad dimension");
  PVQA_REQUIRE(d % 8 == 0, PVQA_ERR_SHAPE, "col_sum: d must be a multiple of 8");
  PVQA_REQUIRE(out, PVQA_ERR_NULL, "col_sum: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)d * sizeof(float), st);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "col_sum: memset: %s", cudaGetErrorString(e));
  }
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(x, PVQA_ERR_NULL, "col_sum: NULL pointer");
  PVQA_REQUIRE(aligned32(x), PVQA_ERR_ALIGN, "col_sum: 32-byte alignment required");
  const int slabs = (int)((d + 255) / 256);
  long long want = ((long long)num_sms() * 4 + slabs - 1) / slabs;      // ~4 CTAs per SM in total
  long long rows_per = (N + want - 1) / want;
  if (rows_per < 32) rows_per = 32;
  rows_per = (rows_per + 7) / 8 * 8;
  dim3 grid((unsigned)slabs, (unsigned)((N + rows_per - 1) / rows_per));
  if (dtype == PVQA_BF16)
    col_sum_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, st>>>((const __nv_bfloat16*)x, out, (int)N, (int)d, (int)rows_per);
  else if (dtype == PVQA_F32)
    col_sum_kernel<float><<<grid, kNormThreads, 0, st>>>((const float*)x, out, (int)N, (int)d, (int)rows_per);
  else
    return fail(PVQA_ERR_DTYPE, "col_sum: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("col_sum");
  return PVQA_OK;
}

// ---- fused pre-norm step of the T5 block: hidden_out = hidden + dropout(update); y = T5LayerNorm(hidden_out) ----
extern "C" int pvqa_add_dropout_rms_fwd(const float* hidden, const void* update, int upd_dtype, const float* weight,
                                        float* hidden_out, void* y, int y_dtype, float* rstd, int64_t N, int64_t d,
                                        float eps, float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "add_dropout_rms_fwd: bad dimension");
  PVQA_REQUIRE(d % 8 == 0 && d <= 1024, PVQA_ERR_SHAPE, "add_dropout_rms_fwd: d=%lld must be a multiple of 8 and <= 1024", (long long)d);
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "add_dropout_rms_fwd: dropout_p must be in [0,1)");
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(hidden && update && weight && hidden_out && y, PVQA_ERR_NULL, "add_dropout_rms_fwd: NULL pointer");
  PVQA_REQUIRE(aligned32(hidden) && aligned32(update) && aligned32(weight) && aligned32(hidden_out) && aligned32(y),
               PVQA_ERR_ALIGN, "add_dropout_rms_fwd: 16-byte alignment required");
  PVQA_REQUIRE(y_dtype == PVQA_BF16 || y_dtype == PVQA_F32, PVQA_ERR_DTYPE, "add_dropout_rms_fwd: bad output dtype");
  uint32_t thr; float sc;
  drop_consts(dropout_p, thr, sc);
  const int warps = kNormThreads / 32, nv = (int)((d + 255) / 256);
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  float* yf = y_dtype == PVQA_F32 ? (float*)y : nullptr;
  bf* yl = y_dtype == PVQA_BF16 ? (bf*)y : nullptr;
  if (upd_dtype == PVQA_BF16)
    launch_add_ln_fwd<float, bf, bf>(nv, grid, st, hidden, (const bf*)update, weight, nullptr, hidden_out, yf, yl, nullptr, rstd, (int)N, (int)d, eps, thr, sc, seed, offset, 1);
  else if (upd_dtype == PVQA_F32)
    launch_add_ln_fwd<float, float, bf>(nv, grid, st, hidden, (const float*)update, weight, nullptr, hidden_out, yf, yl, nullptr, rstd, (int)N, (int)d, eps, thr, sc, seed, offset, 1);
  else
    return fail(PVQA_ERR_DTYPE, "add_dropout_rms_fwd: bad update dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("add_dropout_rms_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_add_dropout_rms_bwd(const void* dy, int y_dtype, const float* d_residual, const float* hidden_out,
                                        const float* weight, const float* rstd, float* d_hidden, void* d_update,
                                        int upd_dtype, float* dweight, int64_t N, int64_t d, float dropout_p,
                                        uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0, PVQA_ERR_SHAPE, "add_dropout_rms_bwd: bad dimension");
  PVQA_REQUIRE(d % 8 == 0 && d <= 1024, PVQA_ERR_SHAPE, "add_dropout_rms_bwd: d must be a multiple of 8 and <= 1024");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "add_dropout_rms_bwd: dropout_p must be in [0,1)");
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(dy && hidden_out && weight && rstd && d_hidden && d_update && dweight, PVQA_ERR_NULL, "add_dropout_rms_bwd: NULL pointer");
  PVQA_REQUIRE(aligned32(dy) && aligned32(d_residual) && aligned32(hidden_out) && aligned32(weight) && aligned32(d_hidden) &&
                   aligned32(d_update), PVQA_ERR_ALIGN, "add_dropout_rms_bwd: 32-byte alignment required");
  PVQA_REQUIRE(y_dtype == PVQA_BF16 || y_dtype == PVQA_F32, PVQA_ERR_DTYPE, "add_dropout_rms_bwd: bad gradient dtype");
  uint32_t thr; float sc;
  drop_consts(dropout_p, thr, sc);
  const int warps = kNormThreads / 32, nv = (int)((d + 255) / 256);
  long long need = (N + warps - 1) / warps, cap = (long long)num_sms() * PVQA_LN_BWD_CTAS;
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)2 * d * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  const float* gf = y_dtype == PVQA_F32 ? (const float*)dy : nullptr;
  const bf* gl = y_dtype == PVQA_BF16 ? (const bf*)dy : nullptr;
  if (upd_dtype == PVQA_BF16)
    launch_add_ln_bwd<bf, bf>(nv, grid, smem, st, gf, gl, d_residual, hidden_out, weight, nullptr, rstd, d_hidden, (bf*)d_update, dweight, nullptr, (int)N, (int)d, thr, sc, seed, offset);
  else if (upd_dtype == PVQA_F32)
    launch_add_ln_bwd<float, bf>(nv, grid, smem, st, gf, gl, d_residual, hidden_out, weight, nullptr, rstd, d_hidden, (float*)d_update, dweight, nullptr, (int)N, (int)d, thr, sc, seed, offset);
  else
    return fail(PVQA_ERR_DTYPE, "add_dropout_rms_bwd: bad update dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("add_dropout_rms_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_cast_rows(const float* src, void* dst, int64_t N, int64_t d, int64_t dst_row_stride, int dst_dtype,
                              void* stream) {
  PVQA_REQUIRE(N >= 0 && d > 0 && d % 8 == 0 && dst_row_stride >= d && dst_row_stride % 8 == 0, PVQA_ERR_SHAPE,
               "cast_rows: d and the destination row stride must be multiples of 8, stride >= d");
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(src && dst, PVQA_ERR_NULL, "cast_rows: NULL pointer");
  PVQA_REQUIRE(aligned32(src) && aligned32(dst), PVQA_ERR_ALIGN, "cast_rows: 32-byte alignment required");
  const long long n8 = (long long)N * (d / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == PVQA_BF16)
    cast_rows_kernel<__nv_bfloat16><<<ew_grid(n8), kNormThreads, 0, st>>>(src, (__nv_bfloat16*)dst, n8, (int)(d / 8), dst_row_stride);
  else if (dst_dtype == PVQA_F32)
    cast_rows_kernel<float><<<ew_grid(n8), kNormThreads, 0, st>>>(src, (float*)dst, n8, (int)(d / 8), dst_row_stride);
  else
    return fail(PVQA_ERR_DTYPE, "cast_rows: bad dtype");
  count_launch();
  PVQA_CHECK_LAUNCH("cast_rows");
  return PVQA_OK;
}
