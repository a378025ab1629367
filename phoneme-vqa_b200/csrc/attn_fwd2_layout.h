// attn_fwd2_layout.h — index arithmetic of the shifted-copy bias staging used by attn_fwd2_kernel, kept free of
// CUDA headers so that tests/test_attn_v2_layout_cpu.py can compile and exhaustively check it with g++.
#pragma once

#if defined(__CUDACC__)
#define PVQA_HD __host__ __device__ __forceinline__
#else
#define PVQA_HD inline
#endif

namespace pvqa_f2 {

constexpr int kRelPadF2 = 128;     // == kRelPad of attn.cu (static_assert there)

// Left padding of the staged vector.  The query row i of a tile reads offsets o + j with o = (Sq-1-i) + pad; rows are
// handed to lanes in order, so the eight lanes of a quarter-warp hold eight consecutive o, the first one
// (i % 8 == 0) at o0 = Sq-1+pad (mod 4).  Choosing pad so that o0 % 4 == 3 makes the eight lanes cover exactly two
// aligned starts x four copies -- the only alignment for which the bank-group argument below holds.
PVQA_HD constexpr int rel_pad(int Sq) { return kRelPadF2 + ((4 - (Sq & 3)) & 3); }

// Elements per shifted copy of the staged bias vector.  copy_k[x] = staged[x + k], k = 0..3, where
// staged[y] = rel_bias[h][y - rel_pad(Sq)] inside [pad, pad + n_rel) and 0 elsewhere.  A thread whose first offset is o reads
// copy_{o&3} from element (o & ~3): 16-byte aligned.  stride % 32 == 8 puts copy k at 16-byte bank group 2k (mod 8)
// relative to copy 0, which makes the eight lanes of a quarter-warp (two aligned starts x four copies) hit eight
// different bank groups.
PVQA_HD constexpr int rel_copy_stride(int Sq, int n_kpad) {
  return ((kRelPadF2 + Sq + n_kpad + 31) / 32) * 32 + 8;
}

// Staging: flat element idx of the [4][cs] array -> index into rel_bias[h][...] (valid iff 0 <= r < n_rel).
PVQA_HD constexpr int rel_copy_source(int idx, int cs, int Sq) {
  return idx - (idx / cs) * cs + (idx / cs) - rel_pad(Sq);
}

// Reading: query row i (absolute) -> element offset (into the [4][cs] array) of the bias of key 0 for that row;
// the bias of key j is at that offset + j.
PVQA_HD constexpr int rel_copy_row_base(int Sq, int i, int cs) {
  return (((Sq - 1 - i) + rel_pad(Sq)) & 3) * cs + (((Sq - 1 - i) + rel_pad(Sq)) & ~3);
}

}  // namespace pvqa_f2

// ---- third-generation forward (attn_fwd3.cuh): same shifted-copy scheme with a 0..3 element left pad.  Rows past the
// end of the sequence (dead rows of the last query tile) reuse the last live row's offset instead of reaching into a
// 128-element pad, which makes the four copies 2 KB smaller — the room the row-max exchange buffer needs to keep two
// CTAs per SM at S = 327.
namespace pvqa_f3 {

PVQA_HD constexpr int rel_pad(int Sq) { return (4 - (Sq & 3)) & 3; }

PVQA_HD constexpr int rel_copy_stride(int Sq, int n_kpad) {
  return ((rel_pad(Sq) + Sq + n_kpad + 31) / 32) * 32 + 8;
}

PVQA_HD constexpr int rel_copy_source(int idx, int cs, int Sq) {
  return idx - (idx / cs) * cs + (idx / cs) - rel_pad(Sq);
}

PVQA_HD constexpr int rel_copy_row_base(int Sq, int i, int cs) {
  return (((Sq - 1 - (i < Sq ? i : Sq - 1)) + rel_pad(Sq)) & 3) * cs +
         (((Sq - 1 - (i < Sq ? i : Sq - 1)) + rel_pad(Sq)) & ~3);
}

}  // namespace pvqa_f3
