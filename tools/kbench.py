"""Per-kernel timing on the B200: CUDA events on the launching stream, L2 flushed between
iterations, reported against MEASURED_PEAKS.json.  Usage:  python tools/kbench.py [embed|tgt|head|attn|all]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


_flush_buf = None


def flush_l2():
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush_buf.zero_()


def time_fn(fn, iters=20, warmup=3, flush=True):
    """median / min milliseconds per call of fn() measured with CUDA events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def embed_mm_bytes(B, S_img, L_ocr, L_q, d, e_tab, e_act):
    """reads B*[(7*L_ocr + L_q)*(d*e_tab + 8) + S_img*d*e_act] + writes B*S*d*e_act (+ mask)."""
    S = S_img + L_ocr + L_q
    return B * ((7 * L_ocr + L_q) * (d * e_tab + 8) + S_img * d * e_act) + B * S * d * e_act


def embed_mm_bwd_bytes(B, S_img, L_ocr, L_q, d, e_act):
    return B * ((L_ocr + L_q) * d * e_act + (7 * L_ocr + L_q) * 8 + 2 * (7 * L_ocr + L_q) * d * 4)


def make_embed_inputs(B=64, S_img=197, L_ocr=100, L_q=30, d=768, V=36096, tab_dtype=torch.float32,
                      act_dtype=torch.bfloat16, seed=1234, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, S_img, d, generator=g).to(act_dtype)
    coords = torch.zeros(B, L_ocr, 6, dtype=torch.long)
    ocr = torch.zeros(B, L_ocr, dtype=torch.long)
    om = torch.zeros(B, L_ocr)
    for b in range(B):
        n = int(torch.randint(20, 100, (1,), generator=g))
        n = min(n, L_ocr - 1)
        x0 = torch.randint(0, 901, (n,), generator=g)
        y0 = torch.randint(0, 901, (n,), generator=g)
        w = torch.randint(1, 101, (n,), generator=g)
        h = torch.randint(1, 101, (n,), generator=g)
        coords[b, :n] = torch.stack([x0, y0, x0 + w, y0 + h, w, h], dim=-1)
        coords[b, n] = 1000
        ocr[b, :n] = torch.randint(3, V, (n,), generator=g)
        ocr[b, n] = 1
        om[b, : n + 1] = 1
    q = torch.zeros(B, L_q, dtype=torch.long)
    qm = torch.zeros(B, L_q)
    for b in range(B):
        n = int(torch.randint(8, L_q + 1, (1,), generator=g))
        q[b, : n - 1] = torch.randint(3, V, (n - 1,), generator=g)
        q[b, n - 1] = 1
        qm[b, :n] = 1
    shared = (torch.randn(V, d, generator=g) * 0.02).to(tab_dtype)
    lay = [(torch.randn(1024, d, generator=g) * 0.02).to(tab_dtype) for _ in range(6)]
    mv = lambda t: t.to(dev)  # noqa: E731
    return mv(img), mv(coords), mv(ocr), mv(q), mv(om), mv(qm), mv(shared), [mv(t) for t in lay]


def bench_embed(B=64):
    from phoneme_vqa_b200 import ops
    pk = peaks()
    res = {}
    for tab_dtype, act_dtype in [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                 (torch.float32, torch.float32)]:
        img, coords, ocr, q, om, qm, shared, lay = make_embed_inputs(B=B, tab_dtype=tab_dtype, act_dtype=act_dtype)
        shared.requires_grad_(True)
        for t in lay:
            t.requires_grad_(True)
        et, ea = shared.element_size(), img.element_size()
        import phoneme_vqa_b200 as pv
        lib = pv.load()
        from phoneme_vqa_b200.ops import _p, _ptr_array, _stream, _dt
        fwd = lambda: ops.embed_multimodal(img, coords, ocr, q, om, qm, shared, lay, out_dtype=act_dtype)  # noqa: E731
        o_buf = torch.empty(B, 327, 768, dtype=act_dtype, device="cuda")
        m_buf = torch.empty(B, 327, dtype=torch.float32, device="cuda")
        larr = _ptr_array(lay)

        def fwd_raw():
            lib.pvqa_embed_mm_fwd(_p(img), _p(coords), _p(ocr), _p(q), _p(om), _p(qm), _p(shared), larr, _p(o_buf),
                                  _p(m_buf), B, 197, 100, 30, 768, shared.shape[0], 1024, _dt(shared.dtype),
                                  _dt(act_dtype), None, _stream())
        med, mn = time_fn(fwd_raw)
        nbytes = embed_mm_bytes(B, 197, 100, 30, 768, et, ea)
        key = f"embed_mm_fwd[tab={str(tab_dtype)[6:]},act={str(act_dtype)[6:]}]"
        res[key] = {"ms_median": med, "ms_min": mn, "alg_MB": nbytes / 1e6, "GBs": nbytes / med / 1e6,
                    "frac_of_" + pk["source"]: nbytes / med / 1e6 / pk["hbm_gbs"]}
        out, _ = fwd()
        go = torch.randn_like(out)
        d_shared = torch.zeros(shared.shape, dtype=torch.float32, device="cuda")
        d_lay = [torch.zeros(t.shape, dtype=torch.float32, device="cuda") for t in lay]
        arr = _ptr_array(d_lay)

        def bwd():
            lib.pvqa_embed_mm_bwd(_p(go), _p(coords), _p(ocr), _p(q), _p(d_shared), arr, B, 197, 100, 30, 768,
                                  shared.shape[0], 1024, _dt(go.dtype), _stream())
        med, mn = time_fn(bwd)
        nbytes = embed_mm_bwd_bytes(B, 197, 100, 30, 768, ea)
        key = f"embed_mm_bwd[act={str(act_dtype)[6:]}]"
        res[key] = {"ms_median": med, "ms_min": mn, "alg_MB": nbytes / 1e6, "GBs": nbytes / med / 1e6,
                    "frac_of_" + pk["source"]: nbytes / med / 1e6 / pk["hbm_gbs"]}
    return res


def bench_norm(N=64 * 327, d=768, iters=10):
    """fused pre-norm glue at the encoder shape: algorithmic bytes / time against the measured HBM peak.
    fwd: read fp32 stream + bf16 update, write stream + bf16 normed; bwd: read dy(bf16) + d_res + z, write d_hidden + d_upd."""
    from phoneme_vqa_b200 import ops
    pk = peaks()
    g = torch.Generator(device="cuda").manual_seed(0)
    res = {}
    w = torch.ones(d, device="cuda", requires_grad=True)
    # rotate over several operand sets so no call finds its inputs in L2 (5 sets x ~260 MB >> 126 MB)
    sets = []
    for _ in range(5):
        h = torch.randn(N, d, device="cuda", generator=g).requires_grad_(True)
        u = torch.randn(N, d, device="cuda", generator=g).bfloat16().requires_grad_(True)
        sets.append((h, u, torch.randn(N, d, device="cuda", generator=g), torch.randn(N, d, device="cuda", generator=g).bfloat16()))
    ops.KernelTimer.reset(True)
    for it in range(3 + iters):
        if it == 3:
            torch.cuda.synchronize()
            ops.KernelTimer.reset(True)
        h, u, g_res, g_y = sets[it % len(sets)]
        ho, y = ops.add_dropout_rms_norm(h, u, w, 1e-6, 0.1, True, torch.bfloat16)
        torch.autograd.backward([ho, y], [g_res, g_y])
        h.grad = u.grad = None
    torch.cuda.synchronize()
    summ = ops.KernelTimer.summary()
    ops.KernelTimer.reset(False)
    alg = {"add_dropout_rms_fwd": N * d * (4 + 2 + 4 + 2), "add_dropout_rms_bwd": N * d * (2 + 4 + 4 + 4 + 2)}
    for kname, (n, tot) in summ.items():
        ms = tot / n
        b = alg.get(kname.split("[")[0])
        if b:
            res[kname] = {"ms": ms, "alg_MB": b / 1e6, "GBs": b / ms / 1e6, "frac_of_hbm": b / ms / 1e6 / pk["hbm_gbs"]}
    return res


def bench_attn(B=64, H=12, S=327, T=127, dropout=0.1, iters=10, only=None):
    """tcgen05 attention kernels at the PhonoLaTr-base shapes; TFLOP/s against the measured bf16 peak.
    flops: fwd 4*B*H*Sq*Sk*D, bwd 10*B*H*Sq*Sk*D (5 GEMMs), causal halves both."""
    import math
    from phoneme_vqa_b200 import ops
    pk = peaks()
    res = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, Sq, Sk, causal, rel, scale in [("enc_self", S, S, False, True, 1.0), ("dec_self", T, T, True, False, 0.125),
                                            ("dec_cross", T, S, False, False, 0.125)]:
        if only and name != only:
            continue
        q = (torch.randn(B, Sq, H, 64, device="cuda", generator=g) * 0.5).bfloat16()
        kv = (torch.randn(B, Sk, 2, H, 64, device="cuda", generator=g) * 0.5).bfloat16()
        k, v = kv[:, :, 0], kv[:, :, 1]
        rb = torch.randn(H, Sq + Sk - 1, device="cuda", generator=g) if rel else None
        ka = torch.zeros(B, Sk, device="cuda")
        go = torch.randn(B, Sq, H, 64, device="cuda", generator=g).bfloat16()
        for p in ((dropout,) if only else (0.0, dropout)):
            drop = (p, 1234, 0)
            fwd = lambda: ops.attention_fwd_raw(q, k, v, scale, rb, ka, causal, drop)  # noqa: E731
            o, lse = fwd()
            dkv = torch.empty_like(kv)
            bwd = lambda: ops.attention_bwd_raw(q, k, v, o, go, lse, scale, rb, ka, causal, dkv[:, :, 0], dkv[:, :, 1],  # noqa: E731
                                                rel, drop)
            ops.KernelTimer.reset(True)
            for _ in range(3):
                fwd(); bwd()
            torch.cuda.synchronize()
            ops.KernelTimer.reset(True)
            for _ in range(iters):
                flush_l2(); fwd(); flush_l2(); bwd()
            torch.cuda.synchronize()
            summ = ops.KernelTimer.summary()
            ops.KernelTimer.reset(False)
            fl = 4.0 * B * H * Sq * Sk * 64 * (0.5 if causal else 1.0)
            for kname, (n, tot) in summ.items():
                ms = tot / n
                f = fl if "fwd" in kname else 2.5 * fl
                res[f"{name}:{kname.split('[')[0]}:p={p}"] = {"ms": ms, "TFLOPs": f / ms / 1e9,
                                                              "frac_of_bf16_peak": f / ms / 1e9 / pk["bf16_tflops"]}
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    out = {"peaks": peaks()}
    if which in ("embed", "all"):
        out.update(bench_embed())
    if which in ("norm", "all"):
        out.update(bench_norm())
    if which in ("attn", "all"):
        out.update(bench_attn())
    if which == "attn_one":         # bench shape, encoder self-attention with dropout, one timed call (for ncu --set full)
        out.update(bench_attn(B=64, iters=1, only="enc_self"))
    if which == "attn_small":       # short run for ncu
        out.update(bench_attn(B=8, iters=1))
    print(json.dumps(out, indent=1))
