// embed.cu — K1 (fused multimodal embedding) and K1' (fused phoneme target embedding + PE)
//
// Both are HBM/L2-bound gather / scatter kernels: no tensor cores.  Forward uses one
// warp per output row with 128-bit loads (7 independent row gathers in flight per lane
// for an OCR row); backward is a run-length-aggregated scatter-add: consecutive tokens
// that hit the same table row (pad tails, eos boxes) are summed in registers and
// flushed with one 16-byte RED per 4 columns, so hot rows do not serialise in L2.
//
// Reference semantics restated (no code shared):
//   core/model/PhonemeLaTr.py:33-44, :219-231      (SpatialModule + concat)
//   PhonoLaTr/modules.py:40-63                      (3-table phoneme embedding, concat)
//   core/model/modules/transformer_utils.py:23-25   (x + PE[:, :T], dropout)
#include "common.cuh"

namespace pvqa {

struct EmbedMMParams {
  const void* img;            // (B, S_img, d) act dtype
  const int64_t* coords;      // (B, L_ocr, 6)
  const int64_t* ocr_ids;     // (B, L_ocr)
  const int64_t* q_ids;       // (B, L_q)
  const float* ocr_mask;      // (B, L_ocr)
  const float* q_mask;        // (B, L_q)
  const void* shared_tab;     // (V, d)
  const void* layout[6];      // (n_pos, d) each
  void* out;                  // (B, S, d)
  float* out_mask;            // (B, S)
  int B, S_img, L_ocr, L_q, d, V, n_pos;
  int32_t* err_flag;
};

constexpr int kFwdThreads = 256;

template <typename TabT, typename ActT>
__global__ void __launch_bounds__(kFwdThreads)
embed_mm_fwd_kernel(const EmbedMMParams p) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = kFwdThreads / 32;
  const int S = p.S_img + p.L_ocr + p.L_q;
  const long long rows = (long long)p.B * S;
  const int chunks = p.d >> 3;
  const TabT* shared_tab = reinterpret_cast<const TabT*>(p.shared_tab);
  ActT* out = reinterpret_cast<ActT*>(p.out);

  for (long long row = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * warps_per_cta) {
    const int b = (int)(row / S);
    const int s = (int)(row - (long long)b * S);
    ActT* orow = out + row * p.d;

    if (s < p.S_img) {
      // ---- image token: streaming copy of the projected ViT row ----
      const ActT* irow = reinterpret_cast<const ActT*>(p.img) + ((long long)b * p.S_img + s) * p.d;
      for (int c = lane; c < chunks; c += 32) {
        f8 a = Vec8<ActT>::load_stream(irow + c * 8);
        Vec8<ActT>::store(orow + c * 8, a);
      }
      if (lane == 0) p.out_mask[row] = 1.0f;
    } else if (s < p.S_img + p.L_ocr) {
      // ---- OCR token: shared[tok] + sum of 6 layout rows ----
      const int l = s - p.S_img;
      const long long tl = (long long)b * p.L_ocr + l;
      long long idx = 0;
      if (lane < 6) idx = p.coords[tl * 6 + lane];
      else if (lane == 6) idx = p.ocr_ids[tl];
      const long long lim = (lane < 6) ? p.n_pos : p.V;
      const bool bad = (lane < 7) && (idx < 0 || idx >= lim);
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0 && p.err_flag) *p.err_flag = 1;
        f8 z;
#pragma unroll
        for (int i = 0; i < 8; ++i) z.v[i] = 0.f;
        for (int c = lane; c < chunks; c += 32) Vec8<ActT>::store(orow + c * 8, z);
        if (lane == 0) p.out_mask[row] = p.ocr_mask[tl];
        continue;
      }
      const TabT* r[7];
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        long long it = __shfl_sync(0xffffffffu, idx, t);
        r[t] = reinterpret_cast<const TabT*>(p.layout[t]) + it * p.d;
      }
      r[6] = shared_tab + __shfl_sync(0xffffffffu, idx, 6) * p.d;
      for (int c = lane; c < chunks; c += 32) {
        f8 a[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) a[t] = Vec8<TabT>::load(r[t] + c * 8);
        f8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // reference association: ((((x0 + y0) + x1) + y1) + w) + h, then ocr + layout
          float lay = a[0].v[i] + a[1].v[i];
          lay += a[2].v[i];
          lay += a[3].v[i];
          lay += a[4].v[i];
          lay += a[5].v[i];
          o.v[i] = a[6].v[i] + lay;
        }
        Vec8<ActT>::store(orow + c * 8, o);
      }
      if (lane == 0) p.out_mask[row] = p.ocr_mask[tl];
    } else {
      // ---- question token: shared[tok] ----
      const int q = s - p.S_img - p.L_ocr;
      const long long tq = (long long)b * p.L_q + q;
      long long tok = 0;
      if (lane == 0) tok = p.q_ids[tq];
      tok = __shfl_sync(0xffffffffu, tok, 0);
      const bool bad = tok < 0 || tok >= p.V;
      if (bad && lane == 0 && p.err_flag) *p.err_flag = 1;
      const TabT* r = shared_tab + (bad ? 0 : tok) * p.d;
      for (int c = lane; c < chunks; c += 32) {
        f8 a = Vec8<TabT>::load(r + c * 8);
        if (bad) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a.v[i] = 0.f;
        }
        Vec8<ActT>::store(orow + c * 8, a);
      }
      if (lane == 0) p.out_mask[row] = p.q_mask[tq];
    }
  }
}

// ---------------------------------------------------------------------------
// K1 backward
// ---------------------------------------------------------------------------
struct EmbedMMBwdParams {
  const void* d_out;          // (B, S, d)
  const int64_t* coords;
  const int64_t* ocr_ids;
  const int64_t* q_ids;
  float* d_shared;            // (V, d) fp32
  float* d_layout[6];         // (n_pos, d) fp32
  int B, S_img, L_ocr, L_q, d, V, n_pos;
  int runs_ocr, runs_q, slabs;
};

constexpr int kRun = 8;           // consecutive tokens aggregated by one warp
constexpr int kBwdThreads = 256;

template <typename ActT> struct Load4;
template <> struct Load4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    uint2 u;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
    float4 r;
    bf16x2_to_f32(u.x, r.x, r.y);
    bf16x2_to_f32(u.y, r.z, r.w);
    return r;
  }
};
template <> struct Load4<float> {
  static __device__ __forceinline__ float4 load(const float* p) {
    uint4 u = ldg16_stream(p);
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  }
};

template <typename ActT>
__global__ void __launch_bounds__(kBwdThreads)
embed_mm_bwd_kernel(const EmbedMMBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int S = p.S_img + p.L_ocr + p.L_q;
  const long long units_ocr = (long long)p.B * p.runs_ocr * p.slabs;
  const long long units = units_ocr + (long long)p.B * p.runs_q * p.slabs;
  const long long total_warps = (long long)gridDim.x * (kBwdThreads / 32);
  const ActT* dout = reinterpret_cast<const ActT*>(p.d_out);

  for (long long u = (long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5); u < units; u += total_warps) {
    if (u < units_ocr) {
      // (b, run, slab) with slab fastest so neighbouring warps share the index loads in L1
      const int slab = (int)(u % p.slabs);
      const long long br = u / p.slabs;
      const int run = (int)(br % p.runs_ocr);
      const int b = (int)(br / p.runs_ocr);
      const int col = slab * 128 + lane * 4;
      const bool active = col < p.d;
      const int t0 = run * kRun;
      const int nt = min(kRun, p.L_ocr - t0);

      float4 g[kRun];
      long long idx[kRun];
#pragma unroll
      for (int i = 0; i < kRun; ++i) {
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        idx[i] = -1;
        if (i < nt) {
          const long long tl = (long long)b * p.L_ocr + t0 + i;
          if (active) g[i] = Load4<ActT>::load(dout + ((long long)b * S + p.S_img + t0 + i) * p.d + col);
          if (lane < 6) idx[i] = p.coords[tl * 6 + lane];
          else if (lane == 6) idx[i] = p.ocr_ids[tl];
        }
      }
      float4 acc[7];
      int cur[7];
#pragma unroll
      for (int t = 0; t < 7; ++t) { cur[t] = -1; acc[t] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int i = 0; i < kRun; ++i) {
        if (i < nt) {
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            long long itl = __shfl_sync(0xffffffffu, idx[i], t);
            const int lim = (t < 6) ? p.n_pos : p.V;
            int it = (itl < 0 || itl >= lim) ? -1 : (int)itl;   // out-of-range: dropped (fwd flagged it)
            if (it != cur[t]) {
              if (cur[t] >= 0 && active) {
                float* dst = ((t < 6) ? p.d_layout[t] : p.d_shared) + (long long)cur[t] * p.d + col;
                red_add_v4(dst, acc[t].x, acc[t].y, acc[t].z, acc[t].w);
              }
              cur[t] = it;
              acc[t] = g[i];
            } else {
              acc[t].x += g[i].x; acc[t].y += g[i].y; acc[t].z += g[i].z; acc[t].w += g[i].w;
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        if (cur[t] >= 0 && active) {
          float* dst = ((t < 6) ? p.d_layout[t] : p.d_shared) + (long long)cur[t] * p.d + col;
          red_add_v4(dst, acc[t].x, acc[t].y, acc[t].z, acc[t].w);
        }
      }
    } else {
      const long long uq = u - units_ocr;
      const int slab = (int)(uq % p.slabs);
      const long long br = uq / p.slabs;
      const int run = (int)(br % p.runs_q);
      const int b = (int)(br / p.runs_q);
      const int col = slab * 128 + lane * 4;
      const bool active = col < p.d;
      const int t0 = run * kRun;
      const int nt = min(kRun, p.L_q - t0);
      // lane i (< kRun) loads token i's id; broadcast below
      long long my = -1;
      if (lane < nt) my = p.q_ids[(long long)b * p.L_q + t0 + lane];
      float4 g[kRun];
#pragma unroll
      for (int i = 0; i < kRun; ++i) {
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nt && active)
          g[i] = Load4<ActT>::load(dout + ((long long)b * S + p.S_img + p.L_ocr + t0 + i) * p.d + col);
      }
      int cur = -1;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kRun; ++i) {
        if (i < nt) {
          long long itl = __shfl_sync(0xffffffffu, my, i);
          int it = (itl < 0 || itl >= p.V) ? -1 : (int)itl;
          if (it != cur) {
            if (cur >= 0 && active)
              red_add_v4(p.d_shared + (long long)cur * p.d + col, acc.x, acc.y, acc.z, acc.w);
            cur = it;
            acc = g[i];
          } else {
            acc.x += g[i].x; acc.y += g[i].y; acc.z += g[i].z; acc.w += g[i].w;
          }
        }
      }
      if (cur >= 0 && active)
        red_add_v4(p.d_shared + (long long)cur * p.d + col, acc.x, acc.y, acc.z, acc.w);
    }
  }
}

// ---------------------------------------------------------------------------
// K1' forward: concat(onset[l0], rhyme[l1], tone[l2]) + PE[t], dropout
// ---------------------------------------------------------------------------
struct EmbedTgtParams {
  const int64_t* labels;      // (B, T, 3)
  const void* tab[3];         // onset (V_o,on), rhyme (V_r,rt), tone (V_t,rt)
  const float* pe;            // (>=T, d)
  void* out;                  // (B, T, d)
  int B, T, d, on_dim, rt_dim;
  int V[3];
  float dropout_p;
  uint64_t seed, offset;
  const unsigned long long* rng_base;
  int32_t* err_flag;
};

// vector path: on_dim % 8 == 0 and rt_dim % 8 == 0 -> one thread per 8 output columns
template <typename TabT, typename ActT>
__global__ void __launch_bounds__(256)
embed_tgt_fwd_vec_kernel(const EmbedTgtParams p) {
  const int chunks = p.d >> 3;
  const long long total = (long long)p.B * p.T * chunks;
  const uint32_t thr16 = (uint32_t)(p.dropout_p * 65536.0f);
  const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  ActT* out = reinterpret_cast<ActT*>(p.out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / chunks;
    const int c = (int)(i - row * chunks);
    const int t = (int)(row % p.T);
    const int col = c * 8;
    int k, off, w;
    if (col < p.on_dim) { k = 0; off = col; w = p.on_dim; }
    else if (col < p.on_dim + p.rt_dim) { k = 1; off = col - p.on_dim; w = p.rt_dim; }
    else { k = 2; off = col - p.on_dim - p.rt_dim; w = p.rt_dim; }
    const long long lab = p.labels[row * 3 + k];
    f8 a;
    if (lab < 0 || lab >= p.V[k]) {
      if (p.err_flag) *p.err_flag = 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] = 0.f;
    } else {
      a = Vec8<TabT>::load(reinterpret_cast<const TabT*>(p.tab[k]) + lab * w + off);
    }
    f8 pe = Vec8<float>::load(p.pe + (long long)t * p.d + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] += pe.v[j];
    if (p.dropout_p > 0.f) {
      const uint32_t m = dropout_keep8(p.seed, p.offset + (p.rng_base ? *p.rng_base : 0ull), (uint64_t)(row * p.d + col), thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] = ((m >> j) & 1u) ? a.v[j] * keep_scale : 0.f;
    }
    Vec8<ActT>::store(out + row * p.d + col, a);
  }
}

// generic path (e.g. d = 512 -> 172/170/170): one thread per element
template <typename TabT, typename ActT>
__global__ void __launch_bounds__(256)
embed_tgt_fwd_scalar_kernel(const EmbedTgtParams p) {
  const long long total = (long long)p.B * p.T * p.d;
  const uint32_t thr16 = (uint32_t)(p.dropout_p * 65536.0f);
  const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  ActT* out = reinterpret_cast<ActT*>(p.out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / p.d;
    const int col = (int)(i - row * p.d);
    const int t = (int)(row % p.T);
    int k, off, w;
    if (col < p.on_dim) { k = 0; off = col; w = p.on_dim; }
    else if (col < p.on_dim + p.rt_dim) { k = 1; off = col - p.on_dim; w = p.rt_dim; }
    else { k = 2; off = col - p.on_dim - p.rt_dim; w = p.rt_dim; }
    const long long lab = p.labels[row * 3 + k];
    float a = 0.f;
    if (lab < 0 || lab >= p.V[k]) { if (p.err_flag) *p.err_flag = 1; }
    else a = to_f32(reinterpret_cast<const TabT*>(p.tab[k])[lab * w + off]);
    a += p.pe[(long long)t * p.d + col];
    if (p.dropout_p > 0.f) {
      const uint64_t e = (uint64_t)i;
      const uint32_t m = dropout_keep8(p.seed, p.offset + (p.rng_base ? *p.rng_base : 0ull), e & ~7ull, thr16);
      a = ((m >> (e & 7)) & 1u) ? a * keep_scale : 0.f;
    }
    out[i] = from_f32<ActT>(a);
  }
}

// ---------------------------------------------------------------------------
// K1' backward: run-length aggregated scatter-add into the 3 sub-tables
// ---------------------------------------------------------------------------
struct EmbedTgtBwdParams {
  const void* d_out;          // (B, T, d)
  const int64_t* labels;      // (B, T, 3)
  float* d_tab[3];
  int B, T, d, on_dim, rt_dim;
  int V[3];
  int runs;                   // ceil(T / kRun)
  int slabs[3];               // column slabs per sub-table
  float dropout_p;
  uint64_t seed, offset;
  const unsigned long long* rng_base;
};

template <typename ActT, int VEC>
__global__ void __launch_bounds__(kBwdThreads)
embed_tgt_bwd_kernel(const EmbedTgtBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int total_slabs = p.slabs[0] + p.slabs[1] + p.slabs[2];
  const long long units = (long long)p.B * p.runs * total_slabs;
  const long long total_warps = (long long)gridDim.x * (kBwdThreads / 32);
  const ActT* dout = reinterpret_cast<const ActT*>(p.d_out);
  const uint32_t thr16 = (uint32_t)(p.dropout_p * 65536.0f);
  const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;

  for (long long u = (long long)blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5); u < units; u += total_warps) {
    int sl = (int)(u % total_slabs);
    const long long br = u / total_slabs;
    const int run = (int)(br % p.runs);
    const int b = (int)(br / p.runs);
    int k = 0;
    if (sl >= p.slabs[0]) { sl -= p.slabs[0]; k = 1; if (sl >= p.slabs[1]) { sl -= p.slabs[1]; k = 2; } }
    const int w = (k == 0) ? p.on_dim : p.rt_dim;
    const int base = (k == 0) ? 0 : (k == 1 ? p.on_dim : p.on_dim + p.rt_dim);
    const int off = sl * 32 * VEC + lane * VEC;     // column within the sub-table
    const bool active = off < w;
    const int t0 = run * kRun;
    const int nt = min(kRun, p.T - t0);

    long long my = -1;
    if (lane < nt) my = p.labels[((long long)b * p.T + t0 + lane) * 3 + k];
    float g[kRun][VEC];
#pragma unroll
    for (int i = 0; i < kRun; ++i) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[i][j] = 0.f;
      if (i < nt && active) {
        const long long e = ((long long)b * p.T + t0 + i) * p.d + base + off;
        if constexpr (VEC == 4) {
          float4 v = Load4<ActT>::load(dout + e);
          g[i][0] = v.x; g[i][1] = v.y; g[i][2] = v.z; g[i][3] = v.w;
        } else {
          g[i][0] = to_f32(dout[e]);
        }
        if (p.dropout_p > 0.f) {
          const uint32_t m = dropout_keep8(p.seed, p.offset + (p.rng_base ? *p.rng_base : 0ull), (uint64_t)e & ~7ull, thr16);
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            g[i][j] = ((m >> (((uint64_t)e + j) & 7)) & 1u) ? g[i][j] * keep_scale : 0.f;
        }
      }
    }
    int cur = -1;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    float* tab = p.d_tab[k];
#pragma unroll
    for (int i = 0; i < kRun; ++i) {
      if (i < nt) {
        long long itl = __shfl_sync(0xffffffffu, my, i);
        int it = (itl < 0 || itl >= p.V[k]) ? -1 : (int)itl;
        if (it != cur) {
          if (cur >= 0 && active) {
            float* dst = tab + (long long)cur * w + off;
            if constexpr (VEC == 4) red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
            else atomicAdd(dst, acc[0]);
          }
          cur = it;
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = g[i][j];
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] += g[i][j];
        }
      }
    }
    if (cur >= 0 && active) {
      float* dst = tab + (long long)cur * w + off;
      if constexpr (VEC == 4) red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
      else atomicAdd(dst, acc[0]);
    }
  }
}

static int grid_for(long long work_items, int per_cta, int ctas_per_sm) {
  long long need = (work_items + per_cta - 1) / per_cta;
  long long cap = (long long)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_embed_mm_fwd(const void* img_feat, const int64_t* coords, const int64_t* ocr_ids,
                                 const int64_t* q_ids, const float* ocr_mask, const float* q_mask,
                                 const void* shared_tab, const void* const* layout_tabs, void* out,
                                 float* out_mask, int64_t B, int64_t S_img, int64_t L_ocr, int64_t L_q,
                                 int64_t d, int64_t V, int64_t n_pos, int tab_dtype, int act_dtype,
                                 int32_t* err_flag, void* stream) {
  PVQA_REQUIRE(B >= 0 && S_img >= 0 && L_ocr >= 0 && L_q >= 0 && d > 0, PVQA_ERR_SHAPE,
               "embed_mm_fwd: negative dimension");
  PVQA_REQUIRE(d % 8 == 0, PVQA_ERR_SHAPE, "embed_mm_fwd: d=%lld must be a multiple of 8", (long long)d);
  PVQA_REQUIRE((tab_dtype == PVQA_F32 || tab_dtype == PVQA_BF16) && (act_dtype == PVQA_F32 || act_dtype == PVQA_BF16),
               PVQA_ERR_DTYPE, "embed_mm_fwd: dtype must be PVQA_F32 or PVQA_BF16");
  const int64_t S = S_img + L_ocr + L_q;
  if (B == 0 || S == 0) return PVQA_OK;
  PVQA_REQUIRE(out && out_mask && shared_tab, PVQA_ERR_NULL, "embed_mm_fwd: out/out_mask/shared_tab is NULL");
  PVQA_REQUIRE(S_img == 0 || img_feat, PVQA_ERR_NULL, "embed_mm_fwd: img_feat is NULL");
  PVQA_REQUIRE(L_ocr == 0 || (coords && ocr_ids && ocr_mask && layout_tabs), PVQA_ERR_NULL,
               "embed_mm_fwd: OCR inputs are NULL but L_ocr > 0");
  PVQA_REQUIRE(L_q == 0 || (q_ids && q_mask), PVQA_ERR_NULL, "embed_mm_fwd: question inputs are NULL but L_q > 0");
  PVQA_REQUIRE(V > 0 && (L_ocr == 0 || n_pos > 0), PVQA_ERR_SHAPE, "embed_mm_fwd: empty table");
  PVQA_REQUIRE(B * S < (1ll << 31) && V < (1ll << 31) && n_pos < (1ll << 31) && d < (1 << 20), PVQA_ERR_SHAPE,
               "embed_mm_fwd: dimension too large");
  EmbedMMParams p{};
  p.img = img_feat; p.coords = coords; p.ocr_ids = ocr_ids; p.q_ids = q_ids;
  p.ocr_mask = ocr_mask; p.q_mask = q_mask; p.shared_tab = shared_tab;
  PVQA_REQUIRE(aligned32(out) && aligned16(shared_tab) && aligned32(img_feat), PVQA_ERR_ALIGN,
               "embed_mm_fwd: pointers must be 16-byte aligned");
  for (int t = 0; t < 6; ++t) {
    p.layout[t] = L_ocr ? layout_tabs[t] : nullptr;
    PVQA_REQUIRE(L_ocr == 0 || (p.layout[t] && aligned16(p.layout[t])), PVQA_ERR_ALIGN,
                 "embed_mm_fwd: layout table %d NULL or misaligned", t);
  }
  p.out = out; p.out_mask = out_mask;
  p.B = (int)B; p.S_img = (int)S_img; p.L_ocr = (int)L_ocr; p.L_q = (int)L_q;
  p.d = (int)d; p.V = (int)V; p.n_pos = (int)n_pos; p.err_flag = err_flag;
  const int grid = grid_for(B * S, kFwdThreads / 32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (tab_dtype == PVQA_F32 && act_dtype == PVQA_F32)
    embed_mm_fwd_kernel<float, float><<<grid, kFwdThreads, 0, st>>>(p);
  else if (tab_dtype == PVQA_F32 && act_dtype == PVQA_BF16)
    embed_mm_fwd_kernel<float, __nv_bfloat16><<<grid, kFwdThreads, 0, st>>>(p);
  else if (tab_dtype == PVQA_BF16 && act_dtype == PVQA_BF16)
    embed_mm_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kFwdThreads, 0, st>>>(p);
  else
    embed_mm_fwd_kernel<__nv_bfloat16, float><<<grid, kFwdThreads, 0, st>>>(p);
  count_launch();
  PVQA_CHECK_LAUNCH("embed_mm_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_embed_mm_bwd(const void* d_out, const int64_t* coords, const int64_t* ocr_ids,
                                 const int64_t* q_ids, float* d_shared, float* const* d_layout_tabs,
                                 int64_t B, int64_t S_img, int64_t L_ocr, int64_t L_q, int64_t d,
                                 int64_t V, int64_t n_pos, int act_dtype, void* stream) {
  PVQA_REQUIRE(B >= 0 && S_img >= 0 && L_ocr >= 0 && L_q >= 0 && d > 0, PVQA_ERR_SHAPE,
               "embed_mm_bwd: negative dimension");
  PVQA_REQUIRE(d % 8 == 0, PVQA_ERR_SHAPE, "embed_mm_bwd: d=%lld must be a multiple of 8", (long long)d);
  PVQA_REQUIRE(act_dtype == PVQA_F32 || act_dtype == PVQA_BF16, PVQA_ERR_DTYPE, "embed_mm_bwd: bad act dtype");
  if (B == 0 || L_ocr + L_q == 0) return PVQA_OK;
  PVQA_REQUIRE(d_out && d_shared, PVQA_ERR_NULL, "embed_mm_bwd: d_out/d_shared is NULL");
  PVQA_REQUIRE(L_ocr == 0 || (coords && ocr_ids && d_layout_tabs), PVQA_ERR_NULL, "embed_mm_bwd: OCR inputs NULL");
  PVQA_REQUIRE(L_q == 0 || q_ids, PVQA_ERR_NULL, "embed_mm_bwd: q_ids NULL");
  PVQA_REQUIRE(aligned32(d_out) && aligned16(d_shared), PVQA_ERR_ALIGN, "embed_mm_bwd: pointers must be 16-byte (d_out: 32-byte) aligned");
  EmbedMMBwdParams p{};
  p.d_out = d_out; p.coords = coords; p.ocr_ids = ocr_ids; p.q_ids = q_ids; p.d_shared = d_shared;
  for (int t = 0; t < 6; ++t) {
    p.d_layout[t] = L_ocr ? d_layout_tabs[t] : nullptr;
    PVQA_REQUIRE(L_ocr == 0 || (p.d_layout[t] && aligned16(p.d_layout[t])), PVQA_ERR_ALIGN,
                 "embed_mm_bwd: grad layout table %d NULL or misaligned", t);
  }
  p.B = (int)B; p.S_img = (int)S_img; p.L_ocr = (int)L_ocr; p.L_q = (int)L_q;
  p.d = (int)d; p.V = (int)V; p.n_pos = (int)n_pos;
  p.runs_ocr = (int)((L_ocr + kRun - 1) / kRun);
  p.runs_q = (int)((L_q + kRun - 1) / kRun);
  p.slabs = (int)((d + 127) / 128);
  const long long units = (long long)B * (p.runs_ocr + p.runs_q) * p.slabs;
  const int grid = grid_for(units, kBwdThreads / 32, 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == PVQA_BF16) embed_mm_bwd_kernel<__nv_bfloat16><<<grid, kBwdThreads, 0, st>>>(p);
  else embed_mm_bwd_kernel<float><<<grid, kBwdThreads, 0, st>>>(p);
  count_launch();
  PVQA_CHECK_LAUNCH("embed_mm_bwd");
  return PVQA_OK;
}

extern "C" int pvqa_embed_tgt_fwd(const int64_t* labels, const void* onset_tab, const void* rhyme_tab,
                                  const void* tone_tab, const float* pe, void* out, int64_t B, int64_t T,
                                  int64_t d, int64_t on_dim, int64_t rt_dim, int64_t V_o, int64_t V_r,
                                  int64_t V_t, int tab_dtype, int act_dtype, float dropout_p, uint64_t seed,
                                  uint64_t offset, int32_t* err_flag, void* stream) {
  PVQA_REQUIRE(B >= 0 && T >= 0 && d > 0 && on_dim > 0 && rt_dim > 0, PVQA_ERR_SHAPE, "embed_tgt_fwd: bad dimension");
  PVQA_REQUIRE(on_dim + 2 * rt_dim == d, PVQA_ERR_SHAPE, "embed_tgt_fwd: on_dim + 2*rt_dim (%lld) != d (%lld)",
               (long long)(on_dim + 2 * rt_dim), (long long)d);
  PVQA_REQUIRE((tab_dtype == PVQA_F32 || tab_dtype == PVQA_BF16) && (act_dtype == PVQA_F32 || act_dtype == PVQA_BF16),
               PVQA_ERR_DTYPE, "embed_tgt_fwd: bad dtype");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "embed_tgt_fwd: dropout_p must be in [0,1)");
  if (B == 0 || T == 0) return PVQA_OK;
  PVQA_REQUIRE(labels && onset_tab && rhyme_tab && tone_tab && pe && out, PVQA_ERR_NULL, "embed_tgt_fwd: NULL pointer");
  PVQA_REQUIRE(V_o > 0 && V_r > 0 && V_t > 0, PVQA_ERR_SHAPE, "embed_tgt_fwd: empty vocabulary");
  EmbedTgtParams p{};
  p.labels = labels; p.tab[0] = onset_tab; p.tab[1] = rhyme_tab; p.tab[2] = tone_tab; p.pe = pe; p.out = out;
  p.B = (int)B; p.T = (int)T; p.d = (int)d; p.on_dim = (int)on_dim; p.rt_dim = (int)rt_dim;
  p.V[0] = (int)V_o; p.V[1] = (int)V_r; p.V[2] = (int)V_t;
  p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.rng_base = g_rng_base; p.err_flag = err_flag;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (on_dim % 8 == 0) && (rt_dim % 8 == 0) && aligned16(onset_tab) && aligned16(rhyme_tab) &&
                   aligned16(tone_tab) && aligned16(pe) && aligned32(out);
  if (vec) {
    const int grid = grid_for(B * T * (d / 8), 256, 8);
#define LAUNCH_TGT_VEC(TT, AT) embed_tgt_fwd_vec_kernel<TT, AT><<<grid, 256, 0, st>>>(p)
    if (tab_dtype == PVQA_F32 && act_dtype == PVQA_F32) LAUNCH_TGT_VEC(float, float);
    else if (tab_dtype == PVQA_F32) LAUNCH_TGT_VEC(float, __nv_bfloat16);
    else if (act_dtype == PVQA_BF16) LAUNCH_TGT_VEC(__nv_bfloat16, __nv_bfloat16);
    else LAUNCH_TGT_VEC(__nv_bfloat16, float);
#undef LAUNCH_TGT_VEC
  } else {
    const int grid = grid_for(B * T * d, 256, 8);
#define LAUNCH_TGT_SC(TT, AT) embed_tgt_fwd_scalar_kernel<TT, AT><<<grid, 256, 0, st>>>(p)
    if (tab_dtype == PVQA_F32 && act_dtype == PVQA_F32) LAUNCH_TGT_SC(float, float);
    else if (tab_dtype == PVQA_F32) LAUNCH_TGT_SC(float, __nv_bfloat16);
    else if (act_dtype == PVQA_BF16) LAUNCH_TGT_SC(__nv_bfloat16, __nv_bfloat16);
    else LAUNCH_TGT_SC(__nv_bfloat16, float);
#undef LAUNCH_TGT_SC
  }
  count_launch();
  PVQA_CHECK_LAUNCH("embed_tgt_fwd");
  return PVQA_OK;
}

extern "C" int pvqa_embed_tgt_bwd(const void* d_out, const int64_t* labels, float* d_onset, float* d_rhyme,
                                  float* d_tone, int64_t B, int64_t T, int64_t d, int64_t on_dim,
                                  int64_t rt_dim, int64_t V_o, int64_t V_r, int64_t V_t, int act_dtype,
                                  float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
  PVQA_REQUIRE(B >= 0 && T >= 0 && d > 0 && on_dim > 0 && rt_dim > 0, PVQA_ERR_SHAPE, "embed_tgt_bwd: bad dimension");
  PVQA_REQUIRE(on_dim + 2 * rt_dim == d, PVQA_ERR_SHAPE, "embed_tgt_bwd: on_dim + 2*rt_dim != d");
  PVQA_REQUIRE(act_dtype == PVQA_F32 || act_dtype == PVQA_BF16, PVQA_ERR_DTYPE, "embed_tgt_bwd: bad dtype");
  PVQA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PVQA_ERR_SHAPE, "embed_tgt_bwd: dropout_p must be in [0,1)");
  if (B == 0 || T == 0) return PVQA_OK;
  PVQA_REQUIRE(d_out && labels && d_onset && d_rhyme && d_tone, PVQA_ERR_NULL, "embed_tgt_bwd: NULL pointer");
  EmbedTgtBwdParams p{};
  p.d_out = d_out; p.labels = labels; p.d_tab[0] = d_onset; p.d_tab[1] = d_rhyme; p.d_tab[2] = d_tone;
  p.B = (int)B; p.T = (int)T; p.d = (int)d; p.on_dim = (int)on_dim; p.rt_dim = (int)rt_dim;
  p.V[0] = (int)V_o; p.V[1] = (int)V_r; p.V[2] = (int)V_t;
  p.runs = (int)((T + kRun - 1) / kRun);
  p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.rng_base = g_rng_base;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (on_dim % 4 == 0) && (rt_dim % 4 == 0) && aligned16(d_out) && aligned16(d_onset) &&
                   aligned16(d_rhyme) && aligned16(d_tone);
  const int vw = vec ? 128 : 32;
  p.slabs[0] = (int)((on_dim + vw - 1) / vw);
  p.slabs[1] = p.slabs[2] = (int)((rt_dim + vw - 1) / vw);
  const long long units = (long long)B * p.runs * (p.slabs[0] + p.slabs[1] + p.slabs[2]);
  const int grid = grid_for(units, kBwdThreads / 32, 4);
  if (vec) {
    if (act_dtype == PVQA_BF16) embed_tgt_bwd_kernel<__nv_bfloat16, 4><<<grid, kBwdThreads, 0, st>>>(p);
    else embed_tgt_bwd_kernel<float, 4><<<grid, kBwdThreads, 0, st>>>(p);
  } else {
    if (act_dtype == PVQA_BF16) embed_tgt_bwd_kernel<__nv_bfloat16, 1><<<grid, kBwdThreads, 0, st>>>(p);
    else embed_tgt_bwd_kernel<float, 1><<<grid, kBwdThreads, 0, st>>>(p);
  }
  count_launch();
  PVQA_CHECK_LAUNCH("embed_tgt_bwd");
  return PVQA_OK;
}
