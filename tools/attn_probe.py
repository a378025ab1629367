"""Staged device probe of the tcgen05 attention kernels: every stage writes its status to gpurun_out/attn_probe.log
BEFORE and AFTER it runs (fsync'ed), so a hang or a crash still tells where.  Stages: forward / backward against fp32
autograd on the device over shapes that walk the kernel's paths (one tile, several tiles, narrow last tile, several
items per CTA with a bias restage, causal, cross, dropout statistics), then CUDA-event timings at the bench shape.
Usage:  python tools/attn_probe.py [parity|timing|all]"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "attn_probe.log"), "a")
T0 = time.time()


def log(msg):
    LOG.write(f"[{time.time() - T0:6.2f}s] {msg}\n")
    LOG.flush()
    os.fsync(LOG.fileno())
    print(msg, flush=True)


import torch  # noqa: E402
from phoneme_vqa_b200 import ops  # noqa: E402
ops._lib.load()
dev = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def reference(q, k, v, scale, rb, ka, causal):
    Sq, Sk = q.shape[1], k.shape[1]
    qf, kf, vf = [t.float().transpose(1, 2).clone().requires_grad_(True) for t in (q, k, v)]
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    relp = None
    if rb is not None:
        relp = rb.clone().requires_grad_(True)
        i = torch.arange(Sq, device=dev)[:, None]
        j = torch.arange(Sk, device=dev)[None, :]
        s = s + relp[:, (j - i + Sq - 1)][None]
    if ka is not None:
        s = s + ka[:, None, None, :]
    if causal:
        s = s + torch.full((Sq, Sk), float("-inf"), device=dev).triu(1)
    o = torch.matmul(torch.softmax(s, -1), vf)
    return o, torch.logsumexp(s, -1), (qf, kf, vf, relp)


def err(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def case(B, H, Sq, Sk, rel, causal, tag, masked=True, bwd=True, grow=False):
    g = torch.Generator().manual_seed(Sq * 7 + Sk + B)
    q = (torch.randn(B, Sq, H, 64, generator=g) * 0.4).bfloat16().to(dev)
    k = torch.randn(B, Sk, H, 64, generator=g) * 0.4
    if grow:        # logits that grow from key tile to key tile: every tile moves the running reference (redo path)
        k = k * (1.0 + 4.0 * (torch.arange(Sk) // 64).float())[None, :, None, None]
    k = k.bfloat16().to(dev)
    v = torch.randn(B, Sk, H, 64, generator=g).bfloat16().to(dev)
    rb = torch.randn(H, Sq + Sk - 1, generator=g).to(dev) if rel else None
    ka = None
    if masked:
        ka = torch.where(torch.rand(B, Sk, generator=g) > 0.2, 0.0, float("-inf"))
        ka[:, 0] = 0.0
        ka = ka.to(dev)
    scale = 1.0 if rel else 1.0 / math.sqrt(64)
    log(f"{tag}: launching forward B={B} H={H} Sq={Sq} Sk={Sk} rel={rel} causal={causal}")
    o, lse = ops.attention_fwd_raw(q, k, v, scale, rb, ka, causal)
    torch.cuda.synchronize()
    ro, rlse, (qf, kf, vf, relp) = reference(q, k, v, scale, rb, ka, causal)
    log(f"{tag}: forward done: rel err o {err(o.transpose(1, 2), ro):.3e}  max|lse - ref| "
        f"{float((lse - rlse).abs().max()):.3e}  nan: {bool(torch.isnan(o.float()).any())}")
    if not bwd:
        return
    go = torch.randn(B, Sq, H, 64, generator=g).bfloat16().to(dev)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    log(f"{tag}: launching backward")
    dq, d_rel, _ = ops.attention_bwd_raw(q, k, v, o, go, lse, scale, rb, ka, causal, dk, dv, rel)
    torch.cuda.synchronize()
    ro.backward(go.float().transpose(1, 2))
    msg = (f"{tag}: backward done: rel err dq {err(dq, qf.grad.transpose(1, 2)):.3e} dk {err(dk, kf.grad.transpose(1, 2)):.3e} "
           f"dv {err(dv, vf.grad.transpose(1, 2)):.3e}")
    if rel:
        msg += f" d_rel {err(d_rel, relp.grad):.3e}"
    log(msg)


def dropout_case(B, H, S, p, tag):
    """V = ones: every output element is the row sum of the dropped probabilities, mean 1, variance from the mask"""
    g = torch.Generator().manual_seed(S)
    q = (torch.randn(B, S, H, 64, generator=g) * 0.3).bfloat16().to(dev)
    k = (torch.randn(B, S, H, 64, generator=g) * 0.3).bfloat16().to(dev)
    v = torch.ones(B, S, H, 64).bfloat16().to(dev)
    log(f"{tag}: launching dropout forward S={S} p={p}")
    o, lse = ops.attention_fwd_raw(q, k, v, 0.125, None, None, False, (p, 1234, 0))
    torch.cuda.synchronize()
    log(f"{tag}: dropout forward done: mean row sum {float(o.float().mean()):.4f} (expect 1), std {float(o.float()[..., 0].std()):.4f}")


if what in ("parity", "all"):
    case(1, 1, 64, 64, False, False, "A one 64x64 tile", masked=False)
    case(1, 1, 128, 128, False, False, "B 128x128 plain")
    case(2, 2, 327, 327, True, False, "C 327 rel (narrow last tile)")
    case(2, 2, 127, 127, False, True, "D 127 causal")
    case(2, 2, 127, 327, False, False, "E cross 127x327")
    case(1, 2, 200, 200, True, True, "F 200 rel causal")
    case(40, 12, 327, 327, True, False, "G 40x12x327 persistent (bias restage)")
    case(3, 1, 1, 5, True, False, "H tiny 1x5")
    case(2, 2, 327, 327, True, False, "J 327 rel, growing logits (reference moves every tile)", grow=True)
    case(2, 2, 200, 200, False, False, "K 200 no mask", masked=False)
    dropout_case(2, 2, 327, 0.1, "I")
    log("parity stages finished")

if what in ("timing", "all"):
    B, H, S = 64, 12, 327
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
        med, mn = timed(lambda: ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop))
        log(f"p={p} forward: median {med:.1f} us, min {mn:.1f} us")
        o, lse = ops.attention_fwd_raw(q, k, v, 1.0, rb, ka, False, drop)
        for far in (0, 91):
            med, mn = timed(lambda: ops.attention_bwd_raw(q, k, v, o, go, lse, 1.0, rb, ka, False, dkv[:, :, 0], dkv[:, :, 1],
                                                          True, drop, rel_far=far))
            log(f"p={p} backward (prep + main + zero-fill) rel_far={far}: median {med:.1f} us, min {mn:.1f} us")
    log("timing done")
