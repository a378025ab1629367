"""Phase timeline of the tcgen05 attention kernels (developer tool, not part of the product path).

`python tools/attn_trace.py build` compiles a private copy of the library with -DPVQA_ATTN_TRACE into
tools/_trace/ (git-ignored, travels with gpurun); `python tools/attn_trace.py [B]` runs the encoder self-attention
shape on the GPU and prints, per traced event, the mean clock64() delta to the previous event over the first 64 CTAs
(thread 0 = a softmax / compute thread, and the issuer thread).
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "tools", "_trace", "libpvqa_trace.so")

# event numbering of the persistent kernels (csrc/attn_fwd.cuh, attn_bwd.cuh): softmax / compute thread 0 | issuer
FWD_EV = {0: "start", 1: "prologue done", 26: "item 0: tiles done | -", 27: "item 0: last PV done | -",
          28: "item 0 stored | -", 29: "item 1 stored | -", 30: "all items done", 31: "end"}
for t in range(3):
    FWD_EV.update({2 + 6 * t: f"t{t} tile start | turn start", 3 + 6 * t: f"t{t} S ready | S(t+1) issued",
                   4 + 6 * t: f"t{t} S in registers | V landed", 5 + 6 * t: f"t{t} bias + max done | P ready",
                   6 + 6 * t: f"t{t} O(t-1) done (rescaled) | PV issued", 7 + 6 * t: f"t{t} P stored | loads issued"})
BWD_EV = {0: "start", 1: "prologue done", 24: "item 0: tiles done | -", 25: "item 0: last GEMMs done | -",
          26: "item 0: last dQ staged | -", 27: "item 0: dK/dV stored | -", 28: "item 0 synced | -",
          29: "item 1 done | -", 30: "all items done", 31: "end"}
for t in range(3):
    BWD_EV.update({2 + 6 * t: f"t{t} tile start | turn start", 3 + 6 * t: f"t{t} S,dP ready | barrier 1 passed",
                   4 + 6 * t: f"t{t} math done | S,dP(t+1) issued", 5 + 6 * t: f"t{t} GEMMs(t-1) done | barrier 3 passed",
                   6 + 6 * t: f"t{t} P,dS stored | GEMMs issued", 7 + 6 * t: f"t{t} dQ(t-1) staged | refilled"})


def build():
    from importlib import import_module
    lib = import_module("phoneme_vqa_b200._lib")
    os.makedirs(os.path.dirname(TRACE_LIB), exist_ok=True)
    cmd = ["nvcc"] + lib.NVCC_FLAGS + ["-DPVQA_ATTN_TRACE", "-o", TRACE_LIB] + [os.path.join(lib.CSRC, s) for s in lib.SOURCES]
    subprocess.run(cmd, check=True)
    print("built", TRACE_LIB)


def report(tr, names, title):
    import numpy as np
    tr = np.asarray(tr, dtype=np.int64).reshape(64, 64)
    print(f"== {title}")
    for who, base in (("thread 0", 0), ("last warp / issuer", 32)):
        ev = tr[:, base:base + 32]
        live = ev[:, 0] > 0
        ev = ev[live]
        print(f"  [{who}] {live.sum()} CTAs traced; total {np.mean(ev[:, 31] - ev[:, 0]):.0f} clk")
        prev = 0
        for e in sorted(names):
            if e == 0 or not (ev[:, e] > 0).all():
                continue
            d = ev[:, e] - ev[:, prev]
            print(f"    {names[e]:34s} +{d.mean():8.0f}  (min {d.min():6d} max {d.max():6d})   @{np.mean(ev[:, e] - ev[:, 0]):8.0f}")
            prev = e


def main():
    import torch
    from importlib import import_module
    lib_mod = import_module("phoneme_vqa_b200._lib")
    lib_mod.LIB_PATH = TRACE_LIB
    from phoneme_vqa_b200 import ops
    lib = lib_mod.load()
    lib.pvqa_debug_attn_trace.restype = ctypes.c_int
    lib.pvqa_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    H, S = 12, 327
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.randn(B, S, H, 64, device="cuda", generator=g) * 0.5).bfloat16()
    kv = (torch.randn(B, S, 2, H, 64, device="cuda", generator=g) * 0.5).bfloat16()
    k, v = kv[:, :, 0], kv[:, :, 1]
    rb = torch.randn(H, 2 * S - 1, device="cuda", generator=g)
    ka = torch.zeros(B, S, device="cuda")
    go = torch.randn(B, S, H, 64, device="cuda", generator=g).bfloat16()
    buf = (ctypes.c_longlong * (64 * 64))()
    for p in (0.0, 0.1):
        drop = (p, 1234, 0)
        for _ in range(2):
            o, lse = ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop)
        torch.cuda.synchronize()
        lib.pvqa_debug_attn_trace(None, 1)
        o, lse = ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop)
        torch.cuda.synchronize()
        lib.pvqa_debug_attn_trace(ctypes.addressof(buf), 1)
        report(list(buf), FWD_EV, f"fwd enc_self B={B} p={p}")
        dkv = torch.empty_like(kv)
        for _ in range(2):
            ops.attention_bwd_raw(q, k, v, o, go, lse, 1.0, rb, ka, False, dkv[:, :, 0], dkv[:, :, 1], True, drop)
        torch.cuda.synchronize()
        lib.pvqa_debug_attn_trace(None, 1)
        ops.attention_bwd_raw(q, k, v, o, go, lse, 1.0, rb, ka, False, dkv[:, :, 0], dkv[:, :, 1], True, drop)
        torch.cuda.synchronize()
        lib.pvqa_debug_attn_trace(ctypes.addressof(buf), 1)
        report(list(buf), BWD_EV, f"bwd enc_self B={B} p={p}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        main()
