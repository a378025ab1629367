// head_tc.cu — K4 on the Blackwell tensor path: shared_lm_head GEMM + column split + the three sub-vocabulary heads
// + 3x cross-entropy in ONE kernel.
//
// reference semantics restated (no code shared):
//   core/model/PhonemeLaTr.py:121-130      h = shared_lm_head(x);  onset / rhyme / tone = Linear(h[:, slice_k])
//   core/executor/PhonemeLaTr_Executor.py:181-190   3x CrossEntropyLoss(ignore_index = pad), summed
//
// One CTA owns 128 rows of x (N = B*T rows in all).  For each head k (its slice of the shared output is 256 wide at
// d = 768: 256 | 256 | 256, core/model/PhonemeLaTr.py:69-70):
//   GEMM 1   D1[128 x 256]  = x[128 x 768] . W_shared[256k : 256k+256, :]^T      tcgen05.mma, TMA-fed 3-stage ring, TMEM
//   epilogue D1 + b_shared -> bf16 h: to shared memory (K-major SW128, the A operand of GEMM 2) and to h_out (the
//            backward recomputes from it)
//   GEMM 2   D2[128 x V_k]  = h_k[128 x 256] . W_k^T                              tcgen05.mma, weights through the same ring
//   epilogue D2 + b_k -> online log-sum-exp over the sub-vocabulary, NLL of the target, lse (N,3); then a second pass
//            over D2 (still in TMEM) writes the UNSCALED logit gradient softmax - onehot as bf16 (rows of ignored
//            targets are zero), so the backward is library GEMMs only: the factor g / count_k — known when the whole
//            grid is done — is folded into their small operands (ops._PhonemeHeadFused)
// GEMM 1 of head k+1 runs under the cross-entropy epilogue of head k.  The logits never exist outside TMEM.
// 4 epilogue warps (thread = row = TMEM lane) + 1 issuer warp (warp-uniform, tc05::elect_one per instruction).
#include "common.cuh"
#include "tc05.cuh"

namespace pvqa {

constexpr int kHRows = 128;                       // rows per CTA (UMMA M)
constexpr int kHW = 256;                          // one head's slice of the shared output
constexpr int kHD = 768;                          // d_model this kernel is specialised for
constexpr int kHKB = 64;                          // K block: 128-byte swizzled rows
constexpr int kHVmax = 192;                       // padded sub-vocabulary (UMMA N of GEMM 2): V_k <= 192
constexpr int kHStages = 3;
constexpr int kHStageA = kHRows * kHKB * 2;       // 16 KB  x block
constexpr int kHStageB = kHW * kHKB * 2;          // 32 KB  weight block (W_shared rows, or a head's W_k block: 24 KB)
constexpr int kHStageBytes = kHStageA + kHStageB;
constexpr int kHOffA2 = kHStages * kHStageBytes;  // 144 KB: bf16 h tile, four [128][64] K-major SW128 sub-tiles = 64 KB
constexpr int kHOffBar = kHOffA2 + kHRows * kHW * 2;
constexpr int kHSmem = 1024 + kHOffBar + 256;
constexpr int kHThreads = 160;
constexpr int kHUnitsPerHead = kHD / kHKB + kHW / kHKB;   // 12 GEMM-1 blocks + 4 GEMM-2 blocks

struct HeadTcParams {
  const int64_t* targets;       // (N,3) with row stride tgt_stride
  long long tgt_stride;
  const float* b_shared;        // (768)
  const float* b_head[3];       // (V_k)
  __nv_bfloat16* h_out;         // (N,768) bf16
  float* loss_sum;              // [3]  (zero-initialised by the launcher)
  int* count;                   // [3]
  float* lse;                   // (N,3)
  __nv_bfloat16* dl[3];         // optional (N, round16(V_k)) bf16: softmax - onehot, zero rows for ignored targets
  int N;
  int V[3];
  long long ignore_index;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(tc05::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc05::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(kHThreads, 1)
phoneme_head_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWs,
                       const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                       const __grid_constant__ CUtensorMap tmW2, const HeadTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kHOffBar);   // [3] a ring stage landed           (TMA)
  uint64_t* bar_empty = bar_full + 3;     // [3] a ring stage was consumed                                   (commit)
  uint64_t* bar_d1 = bar_full + 6;        // D1 of a head complete                                           (commit)
  uint64_t* bar_a2 = bar_full + 7;        // h tile of a head in smem, D1 read                               (128 arrivals)
  uint64_t* bar_d2 = bar_full + 8;        // D2 of a head complete                                           (commit)
  uint64_t* bar_e2 = bar_full + 9;        // D2 of a head read                                               (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == 4;
  const int r0 = blockIdx.x * kHRows;

  if (is_issuer) {
    if (lane == 0) {
      tc05::prefetch_tmap(&tmX); tc05::prefetch_tmap(&tmWs); tc05::prefetch_tmap(&tmW0); tc05::prefetch_tmap(&tmW1);
      tc05::prefetch_tmap(&tmW2);
      for (int x = 0; x < 3; ++x) { tc05::mbar_init(bar_full + x, 1); tc05::mbar_init(bar_empty + x, 1); }
      tc05::mbar_init(bar_d1, 1); tc05::mbar_init(bar_a2, 128); tc05::mbar_init(bar_d2, 1); tc05::mbar_init(bar_e2, 128);
      tc05::fence_barrier_init();
    }
    __syncwarp();
    tc05::tmem_alloc(tmem_slot, 512);     // D1 [0,256)   D2 [256,448)
    tc05::tmem_relinquish();
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  tc05::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (is_issuer) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem0 = tc05::smem_u32(smem);
    const uint32_t idesc1 = tc05::idesc_bf16(kHRows, kHW, 0, 0);
    constexpr int kUnits = 3 * kHUnitsPerHead;
    auto load_unit = [&](int u) {
      const int k = u / kHUnitsPerHead, j = u % kHUnitsPerHead, s = u % kHStages;
      uint8_t* sa = smem + s * kHStageBytes;
      uint8_t* sb = sa + kHStageA;
      if (tc05::elect_one()) {
        if (j < kHD / kHKB) {
          tc05::mbar_expect_tx(bar_full + s, kHStageA + kHStageB);
          tma_load_2d(sa, &tmX, bar_full + s, j * kHKB, r0);
          tma_load_2d(sb, &tmWs, bar_full + s, j * kHKB, k * kHW);
        } else {
          tc05::mbar_expect_tx(bar_full + s, kHVmax * kHKB * 2);
          const CUtensorMap* tw = k == 0 ? &tmW0 : (k == 1 ? &tmW1 : &tmW2);
          tma_load_2d(sb, tw, bar_full + s, (j - kHD / kHKB) * kHKB, 0);
        }
      }
    };
    load_unit(0);
    load_unit(1);
    for (int u = 0; u < kUnits; ++u) {
      const int k = u / kHUnitsPerHead, j = u % kHUnitsPerHead, s = u % kHStages;
      if (u + 2 < kUnits) {
        if (u + 2 >= kHStages) tc05::mbar_wait(bar_empty + (u + 2) % kHStages, ((u + 2) / kHStages - 1) & 1);
        load_unit(u + 2);
      }
      tc05::mbar_wait(bar_full + s, (u / kHStages) & 1);
      const uint64_t ad = tc05::desc_sw128_k(smem0 + s * kHStageBytes);
      const uint64_t bd = tc05::desc_sw128_k(smem0 + s * kHStageBytes + kHStageA);
      if (j < kHD / kHKB) {
        // GEMM 1 block.  (D1 of the previous head was read before its arrive on bar_a2, which this warp waited for
        // ahead of that head's GEMM 2.)
        tc05::tc_fence_after_sync();
        if (tc05::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kHKB / 16; ++ks)
            tc05::mma_bf16_ss(tmem_u, tc05::desc_step(ad, ks * 32), tc05::desc_step(bd, ks * 32), idesc1,
                              (j > 0 || ks > 0) ? 1u : 0u);
          tc05::mma_commit(bar_empty + s);
          if (j == kHD / kHKB - 1) tc05::mma_commit(bar_d1);
        }
      } else {
        const int jj = j - kHD / kHKB;
        if (jj == 0) {
          tc05::mbar_wait(bar_a2, k & 1);                    // h tile written, D1 free
          if (k > 0) tc05::mbar_wait(bar_e2, (k - 1) & 1);   // D2 of the previous head consumed
        }
        tc05::tc_fence_after_sync();
        const int vpad = (p.V[k] + 15) & ~15;
        const uint32_t idesc2 = tc05::idesc_bf16(kHRows, vpad, 0, 0);
        const uint64_t a2 = tc05::desc_sw128_k(smem0 + kHOffA2 + jj * (kHRows * 128));
        if (tc05::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kHKB / 16; ++ks)
            tc05::mma_bf16_ss(tmem_u + kHW, tc05::desc_step(a2, ks * 32), tc05::desc_step(bd, ks * 32), idesc2,
                              (jj > 0 || ks > 0) ? 1u : 0u);
          tc05::mma_commit(bar_empty + s);
          if (jj == kHW / kHKB - 1) tc05::mma_commit(bar_d2);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- 4 epilogue warps: thread = row = TMEM lane ----------------
    const int rowl = warp * 32 + lane;
    const int row = r0 + rowl;
    const bool live = row < p.N;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a2_x = tc05::smem_u32(smem + kHOffA2) + rowl * 128 + ((rowl & 7) << 4);
    for (int k = 0; k < 3; ++k) {
      // ---- epilogue 1: D1 + bias -> bf16 h (smem A operand of GEMM 2, and h_out) ----
      tc05::mbar_wait(bar_d1, k & 1);
      tc05::tc_fence_after_sync();
      const float* bs = p.b_shared + k * kHW;
      __nv_bfloat16* hrow = p.h_out + (long long)(live ? row : 0) * kHD + k * kHW;
#pragma unroll 1
      for (int c = 0; c < kHW / 32; ++c) {
        uint32_t r[32];
        tc05::tmem_ld_32x32(tmem_row + c * 32, r);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + c * 32 + q * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bs + c * 32 + q * 8 + 4));
          u.x = f32x2_to_bf16x2(__uint_as_float(r[q * 8 + 0]) + b0.x, __uint_as_float(r[q * 8 + 1]) + b0.y);
          u.y = f32x2_to_bf16x2(__uint_as_float(r[q * 8 + 2]) + b0.z, __uint_as_float(r[q * 8 + 3]) + b0.w);
          u.z = f32x2_to_bf16x2(__uint_as_float(r[q * 8 + 4]) + b1.x, __uint_as_float(r[q * 8 + 5]) + b1.y);
          u.w = f32x2_to_bf16x2(__uint_as_float(r[q * 8 + 6]) + b1.z, __uint_as_float(r[q * 8 + 7]) + b1.w);
          // sub-tile c / 2 (64 columns), 16-byte chunk (c & 1) * 4 + q, 128-byte swizzle
          const uint32_t addr = (a2_x + (uint32_t)((c >> 1) * (kHRows * 128))) ^ (uint32_t)((((c & 1) * 4 + q)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
          if (live) *reinterpret_cast<uint4*>(hrow + c * 32 + q * 8) = u;
        }
      }
      tc05::fence_proxy_async_smem();
      tc05::tc_fence_before_sync();
      tc05::mbar_arrive(bar_a2);
      // ---- epilogue 2: D2 + bias -> online log-sum-exp over the sub-vocabulary, NLL of the target ----
      const int V = p.V[k];
      const float* bk = p.b_head[k];
      const long long tgt = live ? p.targets[(long long)row * p.tgt_stride + k] : p.ignore_index;
      tc05::mbar_wait(bar_d2, k & 1);
      tc05::tc_fence_after_sync();
      float m = -INFINITY, ssum = 0.f, tgt_logit = 0.f;
      const int nchunk = (V + 31) >> 5;
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        uint32_t r[32];
        tc05::tmem_ld_32x32(tmem_row + kHW + c * 32, r);
        tc05::tmem_ld_wait();
        float l[32];
        float cm = -INFINITY;
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const int col = c * 32 + x;
          l[x] = col < V ? __uint_as_float(r[x]) + __ldg(bk + col) : -INFINITY;
          cm = fmaxf(cm, l[x]);
          if ((long long)col == tgt) tgt_logit = l[x];
        }
        const float mn = fmaxf(m, cm);
        float cs = 0.f;
#pragma unroll
        for (int x = 0; x < 32; ++x) cs += __expf(l[x] - mn);
        ssum = ssum * __expf(m - mn) + cs;
        m = mn;
      }
      const float lse = m + __logf(ssum);
      const bool valid = live && tgt != p.ignore_index;
      if (p.dl[k] != nullptr) {
        // second pass over D2 (still in TMEM): the UNSCALED logit gradient softmax - onehot, bf16, rows padded to 16
        // columns.  The backward scales by g / count_k (known only when the whole grid is done) inside its small GEMMs.
        const int vpad = (V + 15) & ~15;
        __nv_bfloat16* drow = p.dl[k] + (long long)(live ? row : 0) * vpad;
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
          uint32_t r[32];
          tc05::tmem_ld_32x32(tmem_row + kHW + c * 32, r);
          tc05::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col0 = c * 32 + q * 8;
            if (col0 < vpad) {
              float g[8];
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                const int col = col0 + x;
                float pr = 0.f;
                if (valid && col < V) {
                  pr = __expf(__uint_as_float(r[q * 8 + x]) + __ldg(bk + col) - lse);
                  if ((long long)col == tgt) pr -= 1.f;
                }
                g[x] = pr;
              }
              uint4 u;
              u.x = f32x2_to_bf16x2(g[0], g[1]); u.y = f32x2_to_bf16x2(g[2], g[3]);
              u.z = f32x2_to_bf16x2(g[4], g[5]); u.w = f32x2_to_bf16x2(g[6], g[7]);
              if (live) *reinterpret_cast<uint4*>(drow + col0) = u;
            }
          }
        }
      }
      tc05::tc_fence_before_sync();
      tc05::mbar_arrive(bar_e2);
      if (live) p.lse[(long long)row * 3 + k] = lse;
      float nll = valid ? lse - tgt_logit : 0.f;
      int cnt = valid ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        nll += __shfl_xor_sync(0xffffffffu, nll, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      if (lane == 0 && cnt > 0) {
        atomicAdd(p.loss_sum + k, nll);
        atomicAdd(p.count + k, cnt);
      }
    }
  }
  tc05::tc_fence_before_sync();
  __syncthreads();
  if (is_issuer) tc05::tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (rows, cols) bf16 row-major -> 2-D map, box {64 columns, box_rows}, 128-byte swizzle, rows past the end read as zero
static int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int box_rows, const char* what) {
  static EncodeTiledFn2 enc = nullptr;
  if (!enc) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn2>(sym);
  }
  if (!enc) return fail(PVQA_ERR_CUDA, "phoneme_head_fused: cuTensorMapEncodeTiled entry point not available");
  if (!aligned16(ptr) || (cols * 2) % 16) return fail(PVQA_ERR_ALIGN, "phoneme_head_fused: %s must be 16-byte aligned", what);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kHKB, (cuuint32_t)box_rows};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PVQA_ERR_CUDA, "phoneme_head_fused: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return PVQA_OK;
}

}  // namespace pvqa

using namespace pvqa;

extern "C" int pvqa_phoneme_head_fused_fwd(const void* x, const void* W_shared, const float* b_shared,
                                           const int64_t* targets, int64_t tgt_row_stride,
                                           const void* W_onset, const float* b_onset, const void* W_rhyme,
                                           const float* b_rhyme, const void* W_tone, const float* b_tone, void* h_out,
                                           float* loss_sum, int32_t* count, float* lse, void* dl_onset,
                                           void* dl_rhyme, void* dl_tone, int64_t N, int64_t d,
                                           int64_t on_dim, int64_t rt_dim, int64_t V_o, int64_t V_r, int64_t V_t,
                                           int64_t ignore_index, void* stream) {
  PVQA_REQUIRE(d == kHD && on_dim == kHW && rt_dim == kHW, PVQA_ERR_SHAPE,
               "phoneme_head_fused: specialised for d = 768 (256 | 256 | 256 slices); got d %lld, %lld / %lld",
               (long long)d, (long long)on_dim, (long long)rt_dim);
  PVQA_REQUIRE(V_o > 0 && V_r > 0 && V_t > 0 && V_o <= kHVmax && V_r <= kHVmax && V_t <= kHVmax, PVQA_ERR_SHAPE,
               "phoneme_head_fused: sub-vocabularies must have 1..%d entries", kHVmax);
  PVQA_REQUIRE(N >= 0, PVQA_ERR_SHAPE, "phoneme_head_fused: bad N");
  PVQA_REQUIRE(loss_sum && count, PVQA_ERR_NULL, "phoneme_head_fused: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss_sum, 0, 3 * sizeof(float), st);
  cudaMemsetAsync(count, 0, 3 * sizeof(int32_t), st);
  if (N == 0) return PVQA_OK;
  PVQA_REQUIRE(x && W_shared && b_shared && targets && W_onset && b_onset && W_rhyme && b_rhyme && W_tone && b_tone &&
                   h_out && lse,
               PVQA_ERR_NULL, "phoneme_head_fused: NULL pointer");
  PVQA_REQUIRE(aligned16(h_out) && aligned16(b_shared), PVQA_ERR_ALIGN, "phoneme_head_fused: h_out / b_shared must be 16-byte aligned");
  PVQA_REQUIRE((dl_onset != nullptr) == (dl_rhyme != nullptr) && (dl_onset != nullptr) == (dl_tone != nullptr), PVQA_ERR_NULL,
               "phoneme_head_fused: the three logit-gradient outputs come together or not at all");
  PVQA_REQUIRE(aligned16(dl_onset) && aligned16(dl_rhyme) && aligned16(dl_tone), PVQA_ERR_ALIGN,
               "phoneme_head_fused: logit-gradient outputs must be 16-byte aligned");
  CUtensorMap tx, tws, tw0, tw1, tw2;
  int rc;
  if ((rc = make_tmap_2d(&tx, x, N, kHD, kHRows, "x"))) return rc;
  if ((rc = make_tmap_2d(&tws, W_shared, kHD, kHD, kHW, "W_shared"))) return rc;
  if ((rc = make_tmap_2d(&tw0, W_onset, V_o, kHW, kHVmax, "W_onset"))) return rc;
  if ((rc = make_tmap_2d(&tw1, W_rhyme, V_r, kHW, kHVmax, "W_rhyme"))) return rc;
  if ((rc = make_tmap_2d(&tw2, W_tone, V_t, kHW, kHVmax, "W_tone"))) return rc;
  HeadTcParams p{};
  p.targets = targets; p.tgt_stride = tgt_row_stride; p.b_shared = b_shared;
  p.b_head[0] = b_onset; p.b_head[1] = b_rhyme; p.b_head[2] = b_tone;
  p.h_out = reinterpret_cast<__nv_bfloat16*>(h_out); p.loss_sum = loss_sum; p.count = count; p.lse = lse;
  p.dl[0] = reinterpret_cast<__nv_bfloat16*>(dl_onset); p.dl[1] = reinterpret_cast<__nv_bfloat16*>(dl_rhyme);
  p.dl[2] = reinterpret_cast<__nv_bfloat16*>(dl_tone);
  p.N = (int)N; p.V[0] = (int)V_o; p.V[1] = (int)V_r; p.V[2] = (int)V_t; p.ignore_index = ignore_index;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(phoneme_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmem);
    if (e != cudaSuccess) return fail(PVQA_ERR_CUDA, "phoneme_head_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  phoneme_head_tc_kernel<<<(unsigned)((N + kHRows - 1) / kHRows), kHThreads, kHSmem, st>>>(tx, tws, tw0, tw1, tw2, p);
  count_launch();
  PVQA_CHECK_LAUNCH("phoneme_head_fused_fwd");
  return PVQA_OK;
}
