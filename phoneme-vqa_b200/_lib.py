"""ctypes binding of libpvqa_sm100.so (declared in include/pvqa.h).

No pybind11 / torch-extension: the boundary is a plain C ABI.  The
library is built in-tree by ``build()`` (called from ``__graft_entry__.build()``) with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no fallback: if the shared
object is missing every op raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpvqa_sm100.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["api.cu", "embed.cu", "head.cu", "head_tc.cu", "attn.cu", "attn_simt.cu", "norm.cu"]
HEADERS = ["common.cuh", "tc05.cuh", "attn_fwd.cuh", "attn_bwd.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

PVQA_F32, PVQA_BF16 = 0, 1
ABI_VERSION = 11  # must equal PVQA_ABI_VERSION in include/pvqa.h


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE, "pvqa.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into csrc/libpvqa_sm100.so (in-tree).  The translation units are
    compiled concurrently (attn.cu alone instantiates twenty tcgen05 kernels) and linked by one nvcc call."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libpvqa_sm100.so")
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    procs = []
    for src in srcs:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        procs.append((src, obj, subprocess.Popen([nvcc] + compile_flags + ["-c", os.path.join(CSRC, src), "-o", obj],
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, pr in procs:
        out = pr.communicate()[0]
        log.append(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH + ".tmp"] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print("".join(log))
    return LIB_PATH


_lib = None

_i64, _vp, _f = c_int64, c_void_p, c_float
_i64x = lambda n: [_i64] * n  # noqa: E731

_SIGNATURES = {
    "pvqa_abi_version": (c_int, []),
    "pvqa_last_error": (c_char_p, []),
    "pvqa_launch_count": (c_int64, []),
    "pvqa_set_rng_step_counter": (c_int, [_vp]),
    "pvqa_set_attn_bwd_waves": (c_int, [c_int]),
    "pvqa_embed_mm_fwd": (c_int, [_vp] * 7 + [POINTER(c_void_p), _vp, _vp] + _i64x(7) + [c_int, c_int, _vp, _vp]),
    "pvqa_embed_mm_bwd": (c_int, [_vp] * 5 + [POINTER(c_void_p)] + _i64x(7) + [c_int, _vp]),
    "pvqa_embed_tgt_fwd": (c_int, [_vp] * 6 + _i64x(8) + [c_int, c_int, _f, c_uint64, c_uint64, _vp, _vp]),
    "pvqa_embed_tgt_bwd": (c_int, [_vp] * 5 + _i64x(8) + [c_int, _f, c_uint64, c_uint64, _vp]),
    "pvqa_phoneme_head_ce_fwd": (c_int, [_vp, _vp, _i64] + [_vp] * 12 + _i64x(8) + [c_int, c_int, _vp]),
    "pvqa_phoneme_head_ce_bwd": (c_int, [_vp, _vp, _i64] + [_vp] * 12 + _i64x(8) + [c_int, c_int, _vp]),
    "pvqa_phoneme_head_fused_fwd": (c_int, [_vp, _vp, _vp, _vp, _i64] + [_vp] * 6 + [_vp] * 7 + _i64x(8) + [_vp]),
    "pvqa_vocab_ce_grad": (c_int, [_vp, _vp, _i64, _vp, _vp, _vp] + _i64x(3) + [_vp]),
    "pvqa_attn_fwd": (c_int, [_vp] * 7 + _i64x(5) + _i64x(12) + [_f, c_int, _f, c_uint64, c_uint64, _vp, _vp, _i64, _i64, _vp]),
    "pvqa_attn_bwd": (c_int, [_vp] * 13 + _i64x(5) + _i64x(21) + [_f, c_int, _f, c_uint64, c_uint64, _vp, _vp, _vp, _i64, _i64,
                              _i64, _vp]),
    "pvqa_attn_f32_fwd": (c_int, [_vp] * 7 + _i64x(5) + _i64x(12) + [_f, c_int, _f, c_uint64, c_uint64, _vp, _vp, _i64, _i64, _vp]),
    "pvqa_rms_norm_fwd": (c_int, [_vp] * 4 + _i64x(2) + [_f, c_int, c_int, _vp]),
    "pvqa_rms_norm_bwd": (c_int, [_vp] * 7 + _i64x(2) + [c_int, c_int, _vp]),
    "pvqa_residual_dropout_add": (c_int, [_vp] * 3 + [_i64, c_int, _f, c_uint64, c_uint64, _vp]),
    "pvqa_residual_dropout_bwd": (c_int, [_vp] * 2 + [_i64, c_int, _f, c_uint64, c_uint64, _vp]),
    "pvqa_relu_dropout_fwd": (c_int, [_vp] * 2 + [_i64, c_int, _f, c_uint64, c_uint64, _vp]),
    "pvqa_relu_dropout_bwd": (c_int, [_vp] * 3 + [_i64, c_int, _f, _vp]),
    "pvqa_add_dropout_ln_fwd": (c_int, [_vp, c_int, _vp, c_int, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _i64, _i64, _f, _f,
                                        c_uint64, c_uint64, _vp]),
    "pvqa_add_dropout_ln_bwd": (c_int, [_vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, _i64, _i64, _f,
                                        c_uint64, c_uint64, _vp]),
    "pvqa_add_dropout_rms_fwd": (c_int, [_vp, _vp, c_int, _vp, _vp, _vp, c_int, _vp, _i64, _i64, _f, _f, c_uint64, c_uint64, _vp]),
    "pvqa_add_dropout_rms_bwd": (c_int, [_vp, c_int, _vp, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _i64, _i64, _f, c_uint64,
                                         c_uint64, _vp]),
    "pvqa_cast_rows": (c_int, [_vp, _vp, _i64, _i64, _i64, c_int, _vp]),
    "pvqa_col_sum": (c_int, [_vp, _vp, _i64, _i64, c_int, c_int, _vp]),
    "pvqa_attn_f32_bwd": (c_int, [_vp] * 13 + _i64x(5) + _i64x(24) + [_f, c_int, _f, c_uint64, c_uint64, _vp, _vp, _vp, _i64, _i64, _vp]),
}


def register_signature(name, restype, argtypes):
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None and hasattr(_lib, name):
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def load() -> ctypes.CDLL:
    """Return the loaded library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PVQA_LIB_PATH", LIB_PATH)      # A/B builds of the same ABI (tools/); the product uses LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the sm_100a extension has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        if not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    got = lib.pvqa_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libpvqa_sm100.so ABI version {got} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(status: int, what: str = "pvqa") -> None:
    if status != 0:
        msg = load().pvqa_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().pvqa_launch_count())
