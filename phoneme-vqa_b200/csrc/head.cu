// head.cu — K4 (small-vocabulary variant): fused multi-token phoneme head + 3x cross-entropy.
//
// reference semantics restated (no code shared):
//   core/model/PhonemeLaTr.py:124-130                     column split + onset/rhyme/tone Linear heads
//   core/executor/PhonemeLaTr_Executor.py:181-190,263-265 3x CrossEntropyLoss(ignore_index=pad), summed
//
// Shapes are tiny for a GEMM (N=8128 rows, K=256 per head, V = 84/187/7): 1.2 GFLOP against
// 12.5 MB of activations, so the kernel is organised around reading h exactly once and never
// writing logits: one warp owns a row, each lane holds 8 columns of each head's K-slice, the
// 32-way dot-product reductions are done as a 31-shuffle transpose-reduce per block of 32
// vocabulary entries (instead of 5 shuffles per entry), and log-sum-exp is kept online.
// Weights (142 KB in bf16) stay L1/L2 resident.
#include "common.cuh"

namespace pvqa {

struct HeadParams {
  const void* h;              // (N, d) act dtype
  const int64_t* tgt;         // (N, 3) with row stride tgt_stride (elements)
  long long tgt_stride;
  const void* W[3];           // (V_k, w_k) w dtype
  const void* b[3];           // (V_k) w dtype
  float* loss_sum;            // [3]
  int* count;                 // [3]
  float* lse;                 // (N, 3)
  void* logits[3];            // optional (N, V_k) act dtype
  // backward
  const float* grad_loss;     // device scalar
  void* dlogits[3];           // (N, V_k) act dtype
  int N, d, wdim[3], off[3], V[3];
  long long ignore_index;
};

constexpr int kHeadThreads = 256;
constexpr int kMaxChunks = 2;   // per-lane 8-column chunks per head slice: supports w_k <= 512

// sum x[i] over lanes so that lane l ends with the total of x[l] (31 shuffles).
__device__ __forceinline__ float transpose_reduce32(float (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? x[i] : x[i + s];
      const float keep = up ? x[i + s] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

template <typename WT>
__device__ __forceinline__ float dot_chunks(const float (&hv)[kMaxChunks][8], const WT* wrow, int lane, int w) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < w) {
      f8 wv = Vec8<WT>::load(wrow + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(hv[c][j], wv.v[j], acc);
    }
  }
  return acc;
}

// MODE 0: forward (loss sums, counts, lse, optional logits); MODE 1: backward (dlogits)
template <typename WT, typename AT, int MODE>
__global__ void __launch_bounds__(kHeadThreads)
phoneme_head_kernel(const HeadParams p) {
  const int lane = threadIdx.x & 31;
  const int warps = kHeadThreads / 32;
  const AT* h = reinterpret_cast<const AT*>(p.h);
  float loss_acc[3] = {0.f, 0.f, 0.f};
  int cnt_acc[3] = {0, 0, 0};
  float gl = 0.f;
  if (MODE == 1) gl = *p.grad_loss;

  for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < p.N; n += gridDim.x * warps) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int w = p.wdim[k], V = p.V[k];
      const WT* W = reinterpret_cast<const WT*>(p.W[k]);
      const WT* bias = reinterpret_cast<const WT*>(p.b[k]);
      // this lane's columns of the head's K-slice
      float hv[kMaxChunks][8];
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < w) {
          f8 t = Vec8<AT>::load(h + (long long)n * p.d + p.off[k] + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) hv[c][j] = t.v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) hv[c][j] = 0.f;
        }
      }
      const long long tgt = p.tgt[(long long)n * p.tgt_stride + k];
      const bool valid = tgt != p.ignore_index;

      if (MODE == 0) {
        // online log-sum-exp, lane l tracks vocabulary entries v0 + l
        float m = -INFINITY, s = 0.f, tl = 0.f;
        for (int v0 = 0; v0 < V; v0 += 32) {
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (v0 + i < V) ? dot_chunks<WT>(hv, W + (long long)(v0 + i) * w, lane, w) : 0.f;
          float logit = transpose_reduce32(x, lane);
          const int v = v0 + lane;
          if (v < V) {
            logit += to_f32(bias[v]);
            if (p.logits[k]) reinterpret_cast<AT*>(p.logits[k])[(long long)n * V + v] = from_f32<AT>(logit);
            const float mn = fmaxf(m, logit);
            s = s * __expf(m - mn) + __expf(logit - mn);
            m = mn;
            if (v == tgt) tl = logit;
          }
        }
        // merge lanes
        float M = m;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        float S = (m == -INFINITY) ? 0.f : s * __expf(m - M);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          S += __shfl_xor_sync(0xffffffffu, S, o);
          tl += __shfl_xor_sync(0xffffffffu, tl, o);
        }
        const float lse = M + __logf(S);
        if (lane == 0) p.lse[(long long)n * 3 + k] = lse;
        if (valid) { loss_acc[k] += lse - tl; cnt_acc[k] += 1; }
      } else {
        const float lse = p.lse[(long long)n * 3 + k];
        const int cnt = p.count[k];
        const float g = (valid && cnt > 0) ? gl / (float)cnt : 0.f;
        AT* dl = reinterpret_cast<AT*>(p.dlogits[k]) + (long long)n * V;
        for (int v0 = 0; v0 < V; v0 += 32) {
          const int v = v0 + lane;
          if (!valid) {                      // warp-uniform: ignored rows get zero gradient, skip the math
            if (v < V) dl[v] = from_f32<AT>(0.f);
            continue;
          }
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (v0 + i < V) ? dot_chunks<WT>(hv, W + (long long)(v0 + i) * w, lane, w) : 0.f;
          float logit = transpose_reduce32(x, lane);
          if (v < V) {
            logit += to_f32(bias[v]);
            const float prob = __expf(logit - lse);
            dl[v] = from_f32<AT>(g * (prob - (v == tgt ? 1.f : 0.f)));
          }
        }
      }
    }
  }
  if (MODE == 0) {
    // one atomic per warp per head
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (cnt_acc[k]) { atomicAdd(p.loss_sum + k, loss_acc[k]); atomicAdd(p.count + k, cnt_acc[k]); }
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// bf16 tensor-core variant (mma.sync m16n8k16, fp32 accumulate): one warp owns 16 rows; per head the A
// fragments of the 16 x w_k slice stay in registers (w_k <= 256), W_k streams from L1/L2 as B fragments,
// and the C fragments (16 x 8 logits per n-tile) feed an online log-sum-exp — logits still never reach HBM.
// The GEMM is 1.2 GFLOP: far too small for a 128-row tcgen05 tile grid (64 CTAs on 148 SMs), while 16-row
// warp tiles spread it over every SM; the kernel is bounded by reading h once (12.5 MB).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMmaMaxKSteps = 16;     // w_k <= 256

template <int MODE>
__global__ void __launch_bounds__(128)
phoneme_head_mma_kernel(const HeadParams p) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps = 4;
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(p.h);
  const int tiles = (p.N + 15) / 16;
  float loss_acc[3] = {0.f, 0.f, 0.f};
  int cnt_acc[3] = {0, 0, 0};
  float gl = 0.f;
  if (MODE == 1) gl = *p.grad_loss;

  // work item = (16-row tile, head): 3x more warps in flight than one warp per tile
  for (int item = blockIdx.x * warps + (threadIdx.x >> 5); item < tiles * 3; item += gridDim.x * warps) {
    const int tile = item / 3, k = item - tile * 3;
    const int r_lo = tile * 16 + g, r_hi = r_lo + 8;
    const bool ok_lo = r_lo < p.N, ok_hi = r_hi < p.N;
    {
      const int w = p.wdim[k], V = p.V[k], ksteps = w >> 4;
      const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(p.W[k]);
      const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(p.b[k]);
      const long long tg_lo = ok_lo ? p.tgt[(long long)r_lo * p.tgt_stride + k] : p.ignore_index;
      const long long tg_hi = ok_hi ? p.tgt[(long long)r_hi * p.tgt_stride + k] : p.ignore_index;
      const bool v_lo = tg_lo != p.ignore_index, v_hi = tg_hi != p.ignore_index;
      // pad tails make most tiles all-ignored: they contribute no loss and a zero gradient (dlogits is
      // pre-zeroed by the host), so the whole tile is skipped unless the caller wants the logits themselves
      if (!__any_sync(0xffffffffu, v_lo || v_hi) && !(MODE == 0 && p.logits[k])) continue;
      // A fragments of this head's column slice
      uint32_t a[kMmaMaxKSteps][4];
      const __nv_bfloat16* h_lo = h + (long long)r_lo * p.d + p.off[k] + 2 * t;
      const __nv_bfloat16* h_hi = h + (long long)r_hi * p.d + p.off[k] + 2 * t;
#pragma unroll
      for (int ks = 0; ks < kMmaMaxKSteps; ++ks) {
        if (ks < ksteps) {
          a[ks][0] = ok_lo ? *reinterpret_cast<const uint32_t*>(h_lo + ks * 16) : 0u;
          a[ks][1] = ok_hi ? *reinterpret_cast<const uint32_t*>(h_hi + ks * 16) : 0u;
          a[ks][2] = ok_lo ? *reinterpret_cast<const uint32_t*>(h_lo + ks * 16 + 8) : 0u;
          a[ks][3] = ok_hi ? *reinterpret_cast<const uint32_t*>(h_hi + ks * 16 + 8) : 0u;
        }
      }

      float m_lo = -INFINITY, s_lo = 0.f, tl_lo = 0.f, m_hi = -INFINITY, s_hi = 0.f, tl_hi = 0.f;
      float lse_lo = 0.f, lse_hi = 0.f, g_lo = 0.f, g_hi = 0.f;
      if (MODE == 1) {
        const int cnt = p.count[k];
        if (ok_lo) lse_lo = p.lse[(long long)r_lo * 3 + k];
        if (ok_hi) lse_hi = p.lse[(long long)r_hi * 3 + k];
        g_lo = (v_lo && cnt > 0) ? gl / (float)cnt : 0.f;
        g_hi = (v_hi && cnt > 0) ? gl / (float)cnt : 0.f;
      }
      // four n-tiles (32 vocabulary entries) per pass: four independent accumulator chains hide the mma latency
      for (int n0 = 0; n0 < V; n0 += 32) {
        float c[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f; }
        const __nv_bfloat16* wrow[4];
        bool nok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nrow = n0 + q * 8 + g;                      // vocabulary entry this thread feeds as B column
          nok[q] = nrow < V;
          wrow[q] = W + (long long)min(nrow, V - 1) * w + 2 * t;
        }
#pragma unroll
        for (int ks = 0; ks < kMmaMaxKSteps; ++ks) {
          if (ks < ksteps) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t b0 = nok[q] ? __ldg(reinterpret_cast<const uint32_t*>(wrow[q] + ks * 16)) : 0u;
              const uint32_t b1 = nok[q] ? __ldg(reinterpret_cast<const uint32_t*>(wrow[q] + ks * 16 + 8)) : 0u;
              mma_bf16_16816(c[q], a[ks], b0, b1);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = n0 + q * 8 + 2 * t + e;             // this thread's logit columns
            if (col < V) {
              const float bb = __bfloat162float(bias[col]);
              const float l_lo = c[q][e] + bb, l_hi = c[q][2 + e] + bb;
              if (MODE == 0) {
                if (p.logits[k]) {
                  __nv_bfloat16* lg = reinterpret_cast<__nv_bfloat16*>(p.logits[k]);
                  if (ok_lo) lg[(long long)r_lo * V + col] = __float2bfloat16_rn(l_lo);
                  if (ok_hi) lg[(long long)r_hi * V + col] = __float2bfloat16_rn(l_hi);
                }
                float mn = fmaxf(m_lo, l_lo);
                s_lo = s_lo * __expf(m_lo - mn) + __expf(l_lo - mn); m_lo = mn;
                mn = fmaxf(m_hi, l_hi);
                s_hi = s_hi * __expf(m_hi - mn) + __expf(l_hi - mn); m_hi = mn;
                if (col == tg_lo) tl_lo = l_lo;
                if (col == tg_hi) tl_hi = l_hi;
              } else {
                __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(p.dlogits[k]);
                if (ok_lo) dl[(long long)r_lo * V + col] =
                    __float2bfloat16_rn(g_lo * (__expf(l_lo - lse_lo) - (col == tg_lo ? 1.f : 0.f)));
                if (ok_hi) dl[(long long)r_hi * V + col] =
                    __float2bfloat16_rn(g_hi * (__expf(l_hi - lse_hi) - (col == tg_hi ? 1.f : 0.f)));
              }
            }
          }
        }
      }
      if (MODE == 0) {
        // combine the 4 threads of a quad (they hold different columns of the same two rows)
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          float mo = __shfl_xor_sync(0xffffffffu, m_lo, o), so = __shfl_xor_sync(0xffffffffu, s_lo, o);
          float mn = fmaxf(m_lo, mo);
          s_lo = (m_lo == -INFINITY ? 0.f : s_lo * __expf(m_lo - mn)) + (mo == -INFINITY ? 0.f : so * __expf(mo - mn));
          m_lo = mn;
          mo = __shfl_xor_sync(0xffffffffu, m_hi, o); so = __shfl_xor_sync(0xffffffffu, s_hi, o);
          mn = fmaxf(m_hi, mo);
          s_hi = (m_hi == -INFINITY ? 0.f : s_hi * __expf(m_hi - mn)) + (mo == -INFINITY ? 0.f : so * __expf(mo - mn));
          m_hi = mn;
          tl_lo += __shfl_xor_sync(0xffffffffu, tl_lo, o);
          tl_hi += __shfl_xor_sync(0xffffffffu, tl_hi, o);
        }
        if (t == 0) {
          const float l1 = m_lo + __logf(s_lo), l2 = m_hi + __logf(s_hi);
          float dl = 0.f;
          int dc = 0;
          if (ok_lo) { p.lse[(long long)r_lo * 3 + k] = l1; if (v_lo) { dl += l1 - tl_lo; dc += 1; } }
          if (ok_hi) { p.lse[(long long)r_hi * 3 + k] = l2; if (v_hi) { dl += l2 - tl_hi; dc += 1; } }
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) if (kk == k) { loss_acc[kk] += dl; cnt_acc[kk] += dc; }
        }
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float la = loss_acc[k];
      int ca = cnt_acc[k];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        la += __shfl_xor_sync(0xffffffffu, la, o);
        ca += __shfl_xor_sync(0xffffffffu, ca, o);
      }
      if (lane == 0 && ca) { atomicAdd(p.loss_sum + k, la); atomicAdd(p.count + k, ca); }
    }
  }
}

static int validate_head(const char* fn, int64_t N, int64_t d, int64_t on_dim, int64_t rt_dim, int64_t V_o,
                         int64_t V_r, int64_t V_t, int w_dtype, int act_dtype) {
  PVQA_REQUIRE(N >= 0 && d > 0 && on_dim > 0 && rt_dim > 0, PVQA_ERR_SHAPE, "%s: bad dimension", fn);
  PVQA_REQUIRE(on_dim + 2 * rt_dim == d, PVQA_ERR_SHAPE, "%s: on_dim + 2*rt_dim != d", fn);
  PVQA_REQUIRE(on_dim % 8 == 0 && rt_dim % 8 == 0, PVQA_ERR_SHAPE,
               "%s: head slice widths (%lld, %lld) must be multiples of 8", fn, (long long)on_dim, (long long)rt_dim);
  PVQA_REQUIRE(on_dim <= kMaxChunks * 256 && rt_dim <= kMaxChunks * 256, PVQA_ERR_SHAPE,
               "%s: head slice wider than %d", fn, kMaxChunks * 256);
  PVQA_REQUIRE(V_o > 0 && V_r > 0 && V_t > 0, PVQA_ERR_SHAPE, "%s: empty vocabulary", fn);
  PVQA_REQUIRE((w_dtype == PVQA_F32 || w_dtype == PVQA_BF16) && (act_dtype == PVQA_F32 || act_dtype == PVQA_BF16),
               PVQA_ERR_DTYPE, "%s: bad dtype", fn);
  return PVQA_OK;
}

template <int MODE>
static void launch_head(const HeadParams& p, int w_dtype, int act_dtype, cudaStream_t st) {
  const int warps = kHeadThreads / 32;
  long long need = ((long long)p.N + warps - 1) / warps;
  long long cap = (long long)num_sms() * 4;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
  const bool mma_ok = w_dtype == PVQA_BF16 && act_dtype == PVQA_BF16 && p.wdim[0] % 16 == 0 && p.wdim[1] % 16 == 0 &&
                      p.wdim[0] <= 16 * kMmaMaxKSteps && p.wdim[1] <= 16 * kMmaMaxKSteps && p.d % 2 == 0;
  if (mma_ok) {
    const long long tiles = ((long long)p.N + 15) / 16;
    if (MODE == 1) {
      for (int k = 0; k < 3; ++k) cudaMemsetAsync(p.dlogits[k], 0, (size_t)p.N * p.V[k] * 2, st);   // skipped tiles
    } else {
      cudaMemsetAsync(p.lse, 0, (size_t)p.N * 3 * sizeof(float), st);
    }
    long long need_m = (tiles * 3 + 3) / 4, cap_m = (long long)num_sms() * 8;
    phoneme_head_mma_kernel<MODE><<<(int)(need_m < cap_m ? need_m : cap_m), 128, 0, st>>>(p);
  } else if (w_dtype == PVQA_BF16 && act_dtype == PVQA_BF16)
    phoneme_head_kernel<__nv_bfloat16, __nv_bfloat16, MODE><<<grid, kHeadThreads, 0, st>>>(p);
  else if (w_dtype == PVQA_F32 && act_dtype == PVQA_F32)
    phoneme_head_kernel<float, float, MODE><<<grid, kHeadThreads, 0, st>>>(p);
  else if (w_dtype == PVQA_F32)
    phoneme_head_kernel<float, __nv_bfloat16, MODE><<<grid, kHeadThreads, 0, st>>>(p);
  else
    phoneme_head_kernel<__nv_bfloat16, float, MODE><<<grid, kHeadThreads, 0, st>>>(p);
  count_launch();
}

// ---------------------------------------------------------------------------------
// K4 (large-vocabulary variant, LaTr): softmax + cross-entropy + gradient over one CHUNK of logits rows.
// reference: core/model/LaTr.py:83 (lm_head) + core/executor/LaTr_Executor.py:160-163 / base_executor.py:169
// (CrossEntropyLoss(ignore_index=pad)).  The host computes logits chunk-by-chunk with cuBLAS
// (rows x 36096 fp32, ~150 MB, instead of the reference's 1.2 GB logits + 1.2 GB log-softmax + 1.2 GB grad),
// this kernel turns a chunk in place into loss contributions and bf16 dlogits = (softmax - onehot) * inv_count,
// which feed the dh / dW GEMMs of the same chunk.  One CTA per row, two passes (second one mostly from L2).
// ---------------------------------------------------------------------------------
constexpr int kCeThreads = 256;

template <int VEC>
__global__ void __launch_bounds__(kCeThreads)
vocab_ce_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, long long tgt_stride,
                     const float* __restrict__ inv_count, float* __restrict__ loss_sum,
                     __nv_bfloat16* __restrict__ dlogits, int n, int V, long long ignore_index) {
  __shared__ float s_m[kCeThreads / 32], s_s[kCeThreads / 32];
  const int row = blockIdx.x;
  if (row >= n) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (long long)row * V;
  __nv_bfloat16* g = dlogits + (long long)row * V;
  const long long tgt = targets[(long long)row * tgt_stride];
  if (tgt == ignore_index) {                       // ignored row: zero gradient, no loss (CTA-uniform)
    for (int c = tid * VEC; c < V; c += kCeThreads * VEC) {
      if (VEC == 4) *reinterpret_cast<uint2*>(g + c) = make_uint2(0u, 0u);
      else g[c] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  float m = -INFINITY, s = 0.f;
  for (int c = tid * VEC; c < V; c += kCeThreads * VEC) {
    float v[4];
    if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(x + c); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else v[0] = x[c];
    float mx = v[0];
#pragma unroll
    for (int k = 1; k < VEC; ++k) mx = fmaxf(mx, v[k]);
    const float mn = fmaxf(m, mx);
    float add = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) add += __expf(v[k] - mn);
    s = s * __expf(m - mn) + add;
    m = mn;
  }
  // warp then block reduction of (m, s)
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, m, o), so = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, mo);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (mo == -INFINITY ? 0.f : so * __expf(mo - mn));
    m = mn;
  }
  if (lane == 0) { s_m[warp] = m; s_s[warp] = s; }
  __syncthreads();
  float M = s_m[0], S = s_s[0];
#pragma unroll
  for (int w = 1; w < kCeThreads / 32; ++w) {
    const float mn = fmaxf(M, s_m[w]);
    S = S * __expf(M - mn) + s_s[w] * __expf(s_m[w] - mn);
    M = mn;
  }
  const float lse = M + __logf(S);
  const float scale = *inv_count;
  if (tid == 0) atomicAdd(loss_sum, lse - x[tgt]);
  for (int c = tid * VEC; c < V; c += kCeThreads * VEC) {
    float v[4];
    if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(x + c); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else v[0] = x[c];
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = (__expf(v[k] - lse) - ((long long)(c + k) == tgt ? 1.f : 0.f)) * scale;
    if (VEC == 4) {
      uint2 u;
      u.x = f32x2_to_bf16x2(v[0], v[1]);
      u.y = f32x2_to_bf16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(g + c) = u;
    } else {
      g[c] = __float2bfloat16_rn(v[0]);
    }
  }
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_phoneme_head_ce_fwd(const void* h, const int64_t* targets, int64_t tgt_row_stride,
                                        const void* W_onset, const void* b_onset, const void* W_rhyme,
                                        const void* b_rhyme, const void* W_tone, const void* b_tone,
                                        float* loss_sum, int32_t* count, float* lse, void* logits_onset,
                                        void* logits_rhyme, void* logits_tone, int64_t N, int64_t d,
                                        int64_t on_dim, int64_t rt_dim, int64_t V_o, int64_t V_r, int64_t V_t,
                                        int64_t ignore_index, int w_dtype, int act_dtype, void* stream) {
  int rc = validate_head("phoneme_head_ce_fwd", N, d, on_dim, rt_dim, V_o, V_r, V_t, w_dtype, act_dtype);
  if (rc) return rc;
  PVQA_REQUIRE(loss_sum && count, PVQA_ERR_NULL, "phoneme_head_ce_fwd: loss_sum/count NULL");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss_sum, 0, 3 * sizeof(float), st);
  cudaMemsetAsync(count, 0, 3 * sizeof(int32_t), st);
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(h && targets && W_onset && b_onset && W_rhyme && b_rhyme && W_tone && b_tone && lse, PVQA_ERR_NULL,
               "phoneme_head_ce_fwd: NULL pointer");
  PVQA_REQUIRE(aligned16(h) && aligned16(W_onset) && aligned16(W_rhyme) && aligned16(W_tone), PVQA_ERR_ALIGN,
               "phoneme_head_ce_fwd: h / W must be 16-byte aligned");
  HeadParams p{};
  p.h = h; p.tgt = targets; p.tgt_stride = tgt_row_stride;
  p.W[0] = W_onset; p.W[1] = W_rhyme; p.W[2] = W_tone;
  p.b[0] = b_onset; p.b[1] = b_rhyme; p.b[2] = b_tone;
  p.loss_sum = loss_sum; p.count = count; p.lse = lse;
  p.logits[0] = logits_onset; p.logits[1] = logits_rhyme; p.logits[2] = logits_tone;
  p.N = (int)N; p.d = (int)d;
  p.wdim[0] = (int)on_dim; p.wdim[1] = p.wdim[2] = (int)rt_dim;
  p.off[0] = 0; p.off[1] = (int)on_dim; p.off[2] = (int)(on_dim + rt_dim);
  p.V[0] = (int)V_o; p.V[1] = (int)V_r; p.V[2] = (int)V_t;
  p.ignore_index = ignore_index;
  launch_head<0>(p, w_dtype, act_dtype, st);
  PVQA_CHECK_LAUNCH("phoneme_head_ce_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_phoneme_head_ce_bwd(const void* h, const int64_t* targets, int64_t tgt_row_stride,
                                        const void* W_onset, const void* b_onset, const void* W_rhyme,
                                        const void* b_rhyme, const void* W_tone, const void* b_tone,
                                        const float* lse, const int32_t* count, const float* grad_loss,
                                        void* dlogits_onset, void* dlogits_rhyme, void* dlogits_tone, int64_t N,
                                        int64_t d, int64_t on_dim, int64_t rt_dim, int64_t V_o, int64_t V_r,
                                        int64_t V_t, int64_t ignore_index, int w_dtype, int act_dtype,
                                        void* stream) {
  int rc = validate_head("phoneme_head_ce_bwd", N, d, on_dim, rt_dim, V_o, V_r, V_t, w_dtype, act_dtype);
  if (rc) return rc;
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(h && targets && W_onset && b_onset && W_rhyme && b_rhyme && W_tone && b_tone && lse && count &&
                   grad_loss && dlogits_onset && dlogits_rhyme && dlogits_tone,
               PVQA_ERR_NULL, "phoneme_head_ce_bwd: NULL pointer");
  PVQA_REQUIRE(aligned16(h) && aligned16(W_onset) && aligned16(W_rhyme) && aligned16(W_tone), PVQA_ERR_ALIGN,
               "phoneme_head_ce_bwd: h / W must be 16-byte aligned");
  HeadParams p{};
  p.h = h; p.tgt = targets; p.tgt_stride = tgt_row_stride;
  p.W[0] = W_onset; p.W[1] = W_rhyme; p.W[2] = W_tone;
  p.b[0] = b_onset; p.b[1] = b_rhyme; p.b[2] = b_tone;
  p.lse = const_cast<float*>(lse); p.count = const_cast<int*>(count); p.grad_loss = grad_loss;
  p.dlogits[0] = dlogits_onset; p.dlogits[1] = dlogits_rhyme; p.dlogits[2] = dlogits_tone;
  p.N = (int)N; p.d = (int)d;
  p.wdim[0] = (int)on_dim; p.wdim[1] = p.wdim[2] = (int)rt_dim;
  p.off[0] = 0; p.off[1] = (int)on_dim; p.off[2] = (int)(on_dim + rt_dim);
  p.V[0] = (int)V_o; p.V[1] = (int)V_r; p.V[2] = (int)V_t;
  p.ignore_index = ignore_index;
  launch_head<1>(p, w_dtype, act_dtype, (cudaStream_t)stream);
  PVQA_CHECK_LAUNCH("phoneme_head_ce_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_vocab_ce_grad(const float* logits, const int64_t* targets, int64_t tgt_stride,
                                  const float* inv_count, float* loss_sum, void* dlogits_bf16, int64_t n, int64_t V,
                                  int64_t ignore_index, void* stream) {
  PVQA_REQUIRE(n >= 0 && V > 0, PVQA_ERR_SHAPE, "vocab_ce_grad: bad dimension");
  if (n == 0) return PVQA_OK;
  PVQA_REQUIRE(logits && targets && inv_count && loss_sum && dlogits_bf16, PVQA_ERR_NULL, "vocab_ce_grad: NULL pointer");
  PVQA_REQUIRE(n < (1ll << 31) && V < (1ll << 31), PVQA_ERR_SHAPE, "vocab_ce_grad: dimension too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (V % 4 == 0 && aligned16(logits) && (reinterpret_cast<uintptr_t>(dlogits_bf16) & 7) == 0)
    vocab_ce_grad_kernel<4><<<(int)n, kCeThreads, 0, st>>>(logits, targets, tgt_stride, inv_count, loss_sum,
                                                          reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), (int)n, (int)V,
                                                          ignore_index);
  else
    vocab_ce_grad_kernel<1><<<(int)n, kCeThreads, 0, st>>>(logits, targets, tgt_stride, inv_count, loss_sum,
                                                          reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), (int)n, (int)V,
                                                          ignore_index);
  count_launch();
  PVQA_CHECK_LAUNCH("vocab_ce_grad");
  return PVQA_OK;
}
