"""30-second device probe of the two opt-in attention kernels (forward v2, lean backward): every stage writes its
status to gpurun_out/quick_v2_probe.log BEFORE and AFTER it runs, so a hang or a crash still tells where.
Usage (env decides what is exercised):  PVQA_ATTN_BWD_LEAN=1 python tools/quick_v2_probe.py"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick_v2_probe.log"), "a")
T0 = time.time()


def log(msg):
    LOG.write(f"[{time.time() - T0:6.2f}s] {msg}\n")
    LOG.flush()
    os.fsync(LOG.fileno())


LEAN0 = os.environ.get("PVQA_ATTN_BWD_LEAN", "0")
log(f"start, BWD_LEAN={LEAN0}")
import torch  # noqa: E402
log("torch imported")
from phoneme_vqa_b200 import ops  # noqa: E402
ops._lib.load()
log("library loaded")
dev = "cuda:0"


def case(B, H, Sq, Sk, rel, causal, p, tag):
    g = torch.Generator().manual_seed(Sq * 7 + Sk)
    q = (torch.randn(B, Sq, H, 64, generator=g) * 0.4).bfloat16().to(dev)
    k = (torch.randn(B, Sk, H, 64, generator=g) * 0.4).bfloat16().to(dev)
    v = torch.randn(B, Sk, H, 64, generator=g).bfloat16().to(dev)
    rb = torch.randn(H, Sq + Sk - 1, generator=g).to(dev) if rel else None
    ka = torch.where(torch.rand(B, Sk, generator=g) > 0.2, 0.0, float("-inf"))
    ka[:, 0] = 0.0
    ka = ka.to(dev)
    scale = 1.0 if rel else 1.0 / math.sqrt(64)
    drop = (p, 1234, 77) if p > 0 else (0.0, 0, 0)
    ops.ATTN_FWD_V2 = False
    o1, l1 = ops.attention_fwd_raw(q, k, v, scale, rb, ka, causal, drop)
    torch.cuda.synchronize()
    log(f"{tag}: v1 forward done")
    fin = torch.isfinite(l1)
    for ver in ("V2", "V3"):
        setattr(ops, "ATTN_FWD_" + ver, True)
        log(f"{tag}: launching {ver} forward")
        o2, l2 = ops.attention_fwd_raw(q, k, v, scale, rb, ka, causal, drop)
        torch.cuda.synchronize()
        setattr(ops, "ATTN_FWD_" + ver, False)
        log(f"{tag}: {ver} forward done: max|o-o1| = {(o2.float() - o1.float()).abs().max().item():.3e}, "
            f"max|lse-lse1| = {(l2[fin] - l1[fin]).abs().max().item():.3e}, nan in o: {bool(torch.isnan(o2.float()).any())}")
    if p > 0.0:
        # dropout: the mask cannot be reproduced by autograd, so compare the lean backward with the validated one
        go = torch.randn(B, Sq, H, 64, generator=g).bfloat16().to(dev)
        res = {}
        for lean in ("0", "1"):
            os.environ["PVQA_ATTN_BWD_LEAN"] = lean
            dk, dv = torch.zeros_like(k), torch.zeros_like(v)
            log(f"{tag}: launching backward with dropout, lean={lean}")
            dq, d_rel, _ = ops.attention_bwd_raw(q, k, v, o1, go, l1, scale, rb, ka, causal, dk, dv, rel, drop)
            torch.cuda.synchronize()
            res[lean] = (dq, dk, dv) + ((d_rel,) if rel else ())
        diffs = [float((a.float() - b_.float()).abs().max()) for a, b_ in zip(res["0"], res["1"])]
        log(f"{tag}: dropout backward lean vs full: max abs diff dq/dk/dv(/d_rel) = {diffs}")
        os.environ["PVQA_ATTN_BWD_LEAN"] = LEAN0
    if p == 0.0:
        # backward (lean variant when PVQA_ATTN_BWD_LEAN=1 and rel and not causal) against fp32 autograd on the device
        go = torch.randn(B, Sq, H, 64, generator=g).bfloat16().to(dev)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        log(f"{tag}: launching backward")
        dq, d_rel, _ = ops.attention_bwd_raw(q, k, v, o1, go, l1, scale, rb, ka, causal, dk, dv, rel, drop)
        torch.cuda.synchronize()
        qf, kf, vf = [t.float().transpose(1, 2).clone().requires_grad_(True) for t in (q, k, v)]
        s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
        relp = None
        if rel:
            relp = rb.clone().requires_grad_(True)
            i = torch.arange(Sq, device=dev)[:, None]
            j = torch.arange(Sk, device=dev)[None, :]
            s = s + relp[:, (j - i + Sq - 1)][None]
        s = s + ka[:, None, None, :]
        if causal:
            s = s + torch.full((Sq, Sk), float("-inf"), device=dev).triu(1)
        torch.matmul(torch.softmax(s, -1), vf).backward(go.float().transpose(1, 2))
        err = lambda a, b: float((a.float() - b).norm() / (b.norm() + 1e-12))  # noqa: E731
        msg = (f"{tag}: backward done: rel err dq {err(dq, qf.grad.transpose(1, 2)):.3e} dk {err(dk, kf.grad.transpose(1, 2)):.3e} "
               f"dv {err(dv, vf.grad.transpose(1, 2)):.3e}")
        if rel:
            msg += f" d_rel {err(d_rel, relp.grad):.3e}"
        log(msg)


case(1, 1, 128, 128, False, False, 0.0, "A 1x1x128 plain")
case(2, 2, 327, 327, True, False, 0.0, "B 2x2x327 rel")
case(2, 2, 327, 327, True, False, 0.1, "C 2x2x327 rel drop")
case(2, 2, 127, 127, False, True, 0.0, "D 2x2x127 causal")
case(2, 2, 127, 327, False, False, 0.0, "E cross 127x327")
log("all stages finished")
