"""Host-side check of the index arithmetic of the opt-in forward kernel (csrc/attn_fwd2.cuh): the four shifted
copies of the T5 bias vector.  The header with the arithmetic is free of CUDA includes, so it is compiled with g++
and driven through ctypes: every (row, key) must read the bias value the first-generation kernel reads
(`rel_bias[h][j - i + Sq - 1]`), every per-thread start must be 16-byte aligned, and the eight lanes of every
quarter-warp must touch eight different 16-byte bank groups (conflict-free LDS.128)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "phoneme-vqa_b200", "csrc")

SRC = r"""
#include "attn_fwd2_layout.h"
extern "C" int f2_stride(int Sq, int n_kpad) { return pvqa_f2::rel_copy_stride(Sq, n_kpad); }
extern "C" int f2_source(int idx, int cs, int Sq) { return pvqa_f2::rel_copy_source(idx, cs, Sq); }
extern "C" int f2_row_base(int Sq, int i, int cs) { return pvqa_f2::rel_copy_row_base(Sq, i, cs); }
extern "C" int f2_pad() { return pvqa_f2::kRelPadF2; }
extern "C" int f3_stride(int Sq, int n_kpad) { return pvqa_f3::rel_copy_stride(Sq, n_kpad); }
extern "C" int f3_source(int idx, int cs, int Sq) { return pvqa_f3::rel_copy_source(idx, cs, Sq); }
extern "C" int f3_row_base(int Sq, int i, int cs) { return pvqa_f3::rel_copy_row_base(Sq, i, cs); }
"""


@pytest.fixture(scope="module")
def f2(tmp_path_factory):
    d = tmp_path_factory.mktemp("f2")
    src, so = d / "f2.cpp", d / "f2.so"
    src.write_text(SRC)
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{HDR}", str(src), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("Sq,Sk", [(327, 327), (127, 127), (464, 464), (707, 707), (5, 5), (128, 128), (129, 300)])
def test_shifted_copies_return_the_reference_bias(f2, Sq, Sk):
    n_kpad = (Sk + 127) // 128 * 128
    cs = f2.f2_stride(Sq, n_kpad)
    assert cs % 32 == 8 and cs >= f2.f2_pad() + Sq + n_kpad
    n_rel = Sq + Sk - 1
    rel = np.arange(1, n_rel + 1, dtype=np.float64)          # distinct, non-zero
    staged = np.zeros(4 * cs)
    for idx in range(4 * cs):                                 # the kernel's staging loop
        r = f2.f2_source(idx, cs, Sq)
        staged[idx] = rel[r] if 0 <= r < n_rel else 0.0
    n_qpad = (Sq + 127) // 128 * 128
    for i in range(n_qpad):                                   # includes the dead rows past Sq of the last tile
        base = f2.f2_row_base(Sq, i, cs)
        assert base % 4 == 0 and 0 <= base and base + n_kpad <= 4 * cs
        got = staged[base: base + n_kpad]
        j = np.arange(n_kpad)
        r = j - i + Sq - 1
        want = np.where((r >= 0) & (r < n_rel), rel[np.clip(r, 0, n_rel - 1)], 0.0)
        # keys past Sk are masked by the -inf of the staged key vector; only live keys of live rows matter
        if i < Sq:
            assert np.array_equal(got[:Sk], want[:Sk]), (i,)


@pytest.mark.parametrize("Sq", [327, 127, 464, 5, 326, 325, 707])
def test_quarter_warps_are_bank_conflict_free(f2, Sq):
    n_kpad = (Sq + 127) // 128 * 128
    cs = f2.f2_stride(Sq, n_kpad)
    for i0 in range(0, (Sq + 127) // 128 * 128, 128):
        for warp in range(4):
            for quarter in range(4):
                rows = [i0 + warp * 32 + quarter * 8 + l for l in range(8)]
                for step in (0, 4, 28, 96):                    # any 16-byte step inside a chunk: same shift for all lanes
                    groups = {((f2.f2_row_base(Sq, i, cs) + step) // 4) % 8 for i in rows}
                    assert len(groups) == 8, (i0, warp, quarter, step)


# ---- third-generation layout (attn_fwd3.cuh): 0..3 element pad, dead rows clamped to the last live row ----
@pytest.mark.parametrize("Sq,Sk", [(327, 327), (127, 127), (464, 464), (707, 707), (5, 5), (128, 128), (326, 326), (325, 325)])
def test_v3_shifted_copies_return_the_reference_bias(f2, Sq, Sk):
    n_kpad = (Sk + 127) // 128 * 128
    cs = f2.f3_stride(Sq, n_kpad)
    assert cs % 32 == 8
    n_rel = Sq + Sk - 1
    rel = np.arange(1, n_rel + 1, dtype=np.float64)
    staged = np.zeros(4 * cs)
    for idx in range(4 * cs):
        r = f2.f3_source(idx, cs, Sq)
        staged[idx] = rel[r] if 0 <= r < n_rel else 0.0
    for i in range((Sq + 127) // 128 * 128):
        base = f2.f3_row_base(Sq, i, cs)
        assert base % 4 == 0 and 0 <= base and base + n_kpad <= 4 * cs, (i, base)
        if i < Sq:
            j = np.arange(Sk)
            assert np.array_equal(staged[base: base + Sk], rel[j - i + Sq - 1]), (i,)


@pytest.mark.parametrize("Sq", [327, 127, 464, 5, 326, 325, 707])
def test_v3_quarter_warps_are_bank_conflict_free(f2, Sq):
    n_kpad = (Sq + 127) // 128 * 128
    cs = f2.f3_stride(Sq, n_kpad)
    for i0 in range(0, (Sq + 127) // 128 * 128, 128):
        for warp in range(4):
            for quarter in range(4):
                rows = [i0 + warp * 32 + quarter * 8 + l for l in range(8)]
                # lanes past the end share the last live row's address (one broadcast), the others must be distinct
                bases = {f2.f3_row_base(Sq, i, cs) for i in rows}
                groups = {(b // 4) % 8 for b in bases}
                assert len(groups) == len(bases), (i0, warp, quarter)
