"""10-second device timing of the opt-in attention kernels at the bench shape (B=64, H=12, S=327, bias, dropout 0.1):
CUDA events around 5 launches each, after 2 warm-up launches, written to gpurun_out/quick_v2_timing.log."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick_v2_timing.log"), "a")
T0 = time.time()


def log(msg):
    LOG.write(f"[{time.time() - T0:6.2f}s] {msg}\n")
    LOG.flush()
    os.fsync(LOG.fileno())


import torch  # noqa: E402
from phoneme_vqa_b200 import ops  # noqa: E402
ops._lib.load()
dev = "cuda:0"
B, H, S = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 12, 327)))
g = torch.Generator(device=dev).manual_seed(0)
q = (torch.randn(B, S, H, 64, device=dev, generator=g) * 0.5).bfloat16()
kv = (torch.randn(B, S, 2, H, 64, device=dev, generator=g) * 0.5).bfloat16()
k, v = kv[:, :, 0], kv[:, :, 1]
rb = torch.randn(H, 2 * S - 1, device=dev, generator=g)
ka = torch.zeros(B, S, device=dev)
go = torch.randn(B, S, H, 64, device=dev, generator=g).bfloat16()
dkv = torch.empty_like(kv)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


for p in (0.1, 0.0):
    drop = (p, 1234, 0) if p > 0 else (0.0, 0, 0)
    for ver in ("v1", "v2", "v3"):
        ops.ATTN_FWD_V2, ops.ATTN_FWD_V3 = ver == "v2", ver == "v3"
        med, mn = timed(lambda: ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop))
        log(f"p={p} forward {ver}: median {med:.1f} us, min {mn:.1f} us")
    ops.ATTN_FWD_V2 = ops.ATTN_FWD_V3 = False
    o, lse = ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop)
    for lean in ("0", "1"):
        os.environ["PVQA_ATTN_BWD_LEAN"] = lean
        med, mn = timed(lambda: ops.attention_bwd_raw(q, k, v, o, go, lse, 1.0, rb, ka, False, dkv[:, :, 0], dkv[:, :, 1],
                                                      True, drop))
        log(f"p={p} backward (prep + main + zero-fill) {'lean' if lean == '1' else 'full'}: median {med:.1f} us, min {mn:.1f} us")
log("done")
