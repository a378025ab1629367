// common.cuh — shared helpers for libpvqa_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/pvqa.h"

namespace pvqa {

// ---- error plumbing -------------------------------------------------------
char* last_error_buf();                       // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);     // formats into last_error_buf, returns code
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
int num_sms();                                // SM count of the current device (cached)
extern std::atomic<int> g_attn_bwd_waves;     // pvqa_set_attn_bwd_waves: CTAs per SM the attention backward is cut into
// Optional device-resident dropout step counter (pvqa_set_rng_step_counter): every dropout kernel adds
// *g_rng_base to its Philox offset, so a CUDA-graph replay of a whole training step draws fresh masks
// without re-capturing (the host-side offsets are baked into the graph, the counter is not).
extern const unsigned long long* g_rng_base;

#define PVQA_CHECK_LAUNCH(name)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess)                                                         \
      return pvqa::fail(PVQA_ERR_CUDA, "%s: launch failed: %s", name,               \
                        cudaGetErrorString(e__));                                   \
  } while (0)

#define PVQA_REQUIRE(cond, code, ...)                                               \
  do {                                                                              \
    if (!(cond)) return pvqa::fail(code, __VA_ARGS__);                              \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// streaming fp32 rows move as 256-bit accesses (Vec8<float>::load_stream / store): 32-byte alignment
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
inline bool aligned_for(const void* p, int dtype) { return dtype == PVQA_F32 ? aligned32(p) : aligned16(p); }

// ---- 8-element vector load/store in fp32 registers -------------------------
// One "chunk" = 8 consecutive elements: 16 B for bf16, 32 B for fp32.
struct f8 { float v[8]; };

__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// streaming (read-once) 16 B load, bypassing L1 allocation
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg16(void* p, uint4 v) {
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void stg16_stream(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void bf16x2_to_f32(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ f8 load(const __nv_bfloat16* p) {
    uint4 u = ldg16(p);
    f8 r;
    bf16x2_to_f32(u.x, r.v[0], r.v[1]);
    bf16x2_to_f32(u.y, r.v[2], r.v[3]);
    bf16x2_to_f32(u.z, r.v[4], r.v[5]);
    bf16x2_to_f32(u.w, r.v[6], r.v[7]);
    return r;
  }
  static __device__ __forceinline__ f8 load_stream(const __nv_bfloat16* p) {
    uint4 u = ldg16_stream(p);
    f8 r;
    bf16x2_to_f32(u.x, r.v[0], r.v[1]);
    bf16x2_to_f32(u.y, r.v[2], r.v[3]);
    bf16x2_to_f32(u.z, r.v[4], r.v[5]);
    bf16x2_to_f32(u.w, r.v[6], r.v[7]);
    return r;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const f8& a) {
    uint4 u;
    u.x = f32x2_to_bf16x2(a.v[0], a.v[1]);
    u.y = f32x2_to_bf16x2(a.v[2], a.v[3]);
    u.z = f32x2_to_bf16x2(a.v[4], a.v[5]);
    u.w = f32x2_to_bf16x2(a.v[6], a.v[7]);
    stg16_stream(p, u);
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ f8 load(const float* p) {
    uint4 a = ldg16(p), b = ldg16(p + 4);
    f8 r;
    r.v[0] = __uint_as_float(a.x); r.v[1] = __uint_as_float(a.y);
    r.v[2] = __uint_as_float(a.z); r.v[3] = __uint_as_float(a.w);
    r.v[4] = __uint_as_float(b.x); r.v[5] = __uint_as_float(b.y);
    r.v[6] = __uint_as_float(b.z); r.v[7] = __uint_as_float(b.w);
    return r;
  }
  // One 256-bit access per lane (sm_100: LDG/STG.256): a warp covers 1 KB of contiguous sectors.  Two 128-bit
  // accesses per lane would interleave the lanes at a 32-byte stride, every instruction touching only half of each
  // sector it requests, which doubles the L2 -> SM traffic of the un-cached (L1::no_allocate) streaming path.
  static __device__ __forceinline__ f8 load_stream(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                   "=f"(r.v[7]) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ void store(float* p, const f8& a) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]), "f"(a.v[3]), "f"(a.v[4]), "f"(a.v[5]), "f"(a.v[6]),
                    "f"(a.v[7]) : "memory");
  }
};

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}

// pull the line holding p into L2 (no destination register: software pipelining at zero register cost)
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// fp32 vector reduction to global memory (no return): one 16 B RED per 4 floats.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- Philox4x32-10 (counter-based RNG for in-kernel dropout) ----------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// Philox4x32 with R rounds.  R = 7 is the smallest round count that passes BigCrush (Salmon et al., SC'11).
template <int R>
__device__ __forceinline__ uint4 philox4x32_r(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// keep-mask for 8 consecutive elements starting at element index e0 (e0 % 8 == 0).
// One Philox4x32-7 call yields 4x32 bits -> 8x16-bit uniforms; keep iff u16 >= p*65536.  (Forward and backward of every
// fused kernel regenerate their masks through this one function, so the round count is a single decision.)
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t e0, uint32_t thr16) {
  uint64_t c = (e0 >> 3) + offset;
  uint4 r = philox4x32_r<7>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xffffu) >= thr16 ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}

// attention-probability dropout: keep-mask for 16 consecutive keys of one query row.
// group = (row_linear * groups_per_row + j / 16); 8 random bits per element, keep iff byte >= thr8,
// so the drop probability is thr8/256 (p quantised to 1/256; the 1/keep scale uses the quantised value).
__device__ __forceinline__ uint32_t attn_dropout_keep16(uint64_t seed, uint64_t offset, uint64_t group, uint32_t thr8) {
  const uint64_t c = offset + group;
  uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5A17u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) m |= (((w[i] >> (8 * bb)) & 0xffu) >= thr8 ? 1u : 0u) << (4 * i + bb);
  }
  return m;
}

}  // namespace pvqa
