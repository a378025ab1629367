"""Fused glue kernels (RMS norm, residual+dropout, relu+dropout, shadow-weight linear) vs torch math."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("N,d,ydt", [(37, 768, torch.bfloat16), (5, 192, torch.float32), (4096, 1024, torch.bfloat16),
                                     (3, 512, torch.float32)])
def test_rms_norm_fwd_bwd(N, d, ydt):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(N)
    x = torch.randn(N, d, generator=g).to(DEV).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV).requires_grad_(True)
    go = torch.randn(N, d, generator=g).to(DEV).to(ydt)
    y = ops.rms_norm(x, w, 1e-6, ydt)
    y.backward(go)
    xr = x.detach().clone().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    var = xr.pow(2).mean(-1, keepdim=True)
    yr = wr * (xr * torch.rsqrt(var + 1e-6))          # HF T5LayerNorm
    yr.backward(go.float())
    tol = 1e-5 if ydt == torch.float32 else 1e-2
    torch.testing.assert_close(y.float(), yr.detach(), rtol=tol, atol=tol)
    torch.testing.assert_close(x.grad, xr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w.grad, wr.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("udt", [torch.float32, torch.bfloat16])
def test_residual_dropout_add(udt):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(0)
    h = torch.randn(8, 50, 768, generator=g).to(DEV).requires_grad_(True)
    u = (torch.randn(8, 50, 768, generator=g) + 3.0).to(DEV).to(udt).requires_grad_(True)
    out = ops.residual_dropout_add(h, u, 0.1, training=False)
    torch.testing.assert_close(out, h.detach() + u.detach().float())
    ops.manual_seed(5)
    out = ops.residual_dropout_add(h, u, 0.1, training=True)
    delta = out.detach() - h.detach()
    kept = delta != 0
    assert abs(kept.float().mean().item() - 0.9) < 5e-3
    torch.testing.assert_close(delta[kept], (u.detach().float() / 0.9)[kept], rtol=2e-3, atol=2e-3)
    out.sum().backward()
    torch.testing.assert_close(h.grad, torch.ones_like(h))
    torch.testing.assert_close(u.grad.float(), kept.float() / 0.9, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_relu_dropout(dt):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 3072, generator=g).to(DEV).to(dt).requires_grad_(True)
    y = ops.relu_dropout(x, 0.1, training=False)
    assert torch.equal(y, torch.relu(x.detach()))
    ops.manual_seed(9)
    y = ops.relu_dropout(x, 0.1, training=True)
    pos = x.detach() > 0
    kept = y != 0
    assert not (kept & ~pos).any()
    assert abs(kept[pos].float().mean().item() - 0.9) < 1e-2
    torch.testing.assert_close(y[kept].float(), (x.detach().float() / 0.9)[kept], rtol=1e-2, atol=1e-2)
    go = torch.randn(64, 3072, generator=g).to(DEV).to(dt)
    y.backward(go)
    torch.testing.assert_close(x.grad.float(), (go.float() / 0.9) * kept.float(), rtol=1e-2, atol=1e-2)


def test_shadow_weight_linear_fp32_grads_and_refresh():
    from phoneme_vqa_b200 import modules as M
    torch.manual_seed(0)
    lin_q, lin_k = torch.nn.Linear(64, 32).to(DEV), torch.nn.Linear(64, 48).to(DEV)
    x = torch.randn(10, 64, device=DEV).bfloat16().requires_grad_(True)
    y = M._lin_multi(x, [lin_q.weight, lin_k.weight], [lin_q.bias, lin_k.bias])
    assert y.dtype == torch.bfloat16 and y.shape == (10, 80)
    ref = torch.nn.functional.linear(x.float(), torch.cat([lin_q.weight, lin_k.weight]).bfloat16().float(),
                                     torch.cat([lin_q.bias, lin_k.bias]).bfloat16().float())
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)
    go = torch.randn_like(y)
    y.backward(go)
    assert lin_q.weight.grad.dtype == torch.float32
    torch.testing.assert_close(lin_q.weight.grad, (go.float().t() @ x.detach().float())[:32], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(lin_k.bias.grad, go.float().sum(0)[32:], rtol=1e-3, atol=1e-3)
    # an optimizer-style in-place update must invalidate the shadow
    with torch.no_grad():
        lin_q.weight.add_(1.0)
    y2 = M._lin_multi(x, [lin_q.weight, lin_k.weight], [lin_q.bias, lin_k.bias])
    assert (y2[:, :32].float() - y[:, :32].float()).abs().max() > 0.1


def test_rms_norm_with_fused_residual_gradient():
    """(normed, x) variant: d x = d_residual + d(norm branch) from one kernel == plain autograd sum."""
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(6, 50, 768, generator=g).to(DEV).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(768, generator=g)).to(DEV).requires_grad_(True)
    proj = torch.randn(768, 768, generator=g).to(DEV) * 0.03
    normed, xp = ops.rms_norm_residual(x, w, 1e-6, torch.bfloat16)
    out = xp + (normed.float() @ proj)
    go = torch.randn(6, 50, 768, generator=g).to(DEV)
    out.backward(go)
    xr = x.detach().clone().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    nr = ops.rms_norm(xr, wr, 1e-6, torch.bfloat16)
    (xr + (nr.float() @ proj)).backward(go)
    torch.testing.assert_close(x.grad, xr.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(w.grad, wr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("N,d,udt,p", [(37, 768, torch.bfloat16, 0.1), (8128, 768, torch.bfloat16, 0.1),
                                       (5, 192, torch.float32, 0.0), (300, 1024, torch.float32, 0.1),
                                       (9, 8, torch.bfloat16, 0.0)])
def test_add_dropout_layer_norm_matches_unfused_pair(N, d, udt, p):
    """fused LayerNorm(hidden + dropout(update)) == residual_dropout_add (same Philox mask) + torch layer_norm,
    forward and backward, including the bf16 side output and its gradient."""
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(N + d)
    h0 = torch.randn(N, d, generator=g).to(DEV)
    u0 = (torch.randn(N, d, generator=g) * 2).to(DEV).to(udt)
    w0 = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV)
    b0 = (0.1 * torch.randn(d, generator=g)).to(DEV)
    gy = torch.randn(N, d, generator=g).to(DEV)
    gy_lp = torch.randn(N, d, generator=g).to(DEV).bfloat16()
    outs = []
    for fused in (True, False):
        h, u, w, b = (t.clone().requires_grad_(True) for t in (h0, u0, w0, b0))
        ops.manual_seed(11)
        if fused:
            y, y_lp = ops.add_dropout_layer_norm(h, u, w, b, 1e-5, p, training=True, want_lp=True)
            assert y_lp.dtype == torch.bfloat16
        else:
            z = ops.residual_dropout_add(h, u, p, training=True)
            y = torch.nn.functional.layer_norm(z, (d,), w, b, 1e-5)
            y_lp = y.bfloat16()
        torch.autograd.backward([y, y_lp], [gy, gy_lp])
        outs.append((y.detach(), y_lp.detach().float(), h.grad, u.grad.float(), w.grad, b.grad))
    names = ("y", "y_lp", "d_hidden", "d_update", "d_gamma", "d_beta")
    lp_u = udt == torch.bfloat16
    tols = {"y": 2e-5, "y_lp": 2e-2, "d_hidden": 2e-4, "d_update": 2e-2 if lp_u else 2e-4, "d_gamma": 2e-3, "d_beta": 2e-3}
    for n, a, r in zip(names, outs[0], outs[1]):
        torch.testing.assert_close(a, r, rtol=tols[n], atol=tols[n] * max(1.0, float(r.abs().max()) if "gamma" in n or "beta" in n else 1.0),
                                   msg=lambda m, n=n: f"{n}: {m}")


def test_add_dropout_layer_norm_eval_and_single_consumer():
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(3)
    h = torch.randn(6, 10, 768, generator=g).to(DEV).requires_grad_(True)
    u = torch.randn(6, 10, 768, generator=g).to(DEV).bfloat16().requires_grad_(True)
    w = torch.ones(768, device=DEV, requires_grad=True)
    b = torch.zeros(768, device=DEV, requires_grad=True)
    y, y_lp = ops.add_dropout_layer_norm(h, u, w, b, 1e-5, 0.3, training=False, want_lp=False)
    assert y_lp is None
    ref = torch.nn.functional.layer_norm(h.detach() + u.detach().float(), (768,), w.detach(), b.detach(), 1e-5)
    torch.testing.assert_close(y, ref, rtol=2e-5, atol=2e-5)
    # only the bf16 copy is consumed downstream: the fp32 gradient arrives as None
    y, y_lp = ops.add_dropout_layer_norm(h, u, w, b, 1e-5, 0.0, training=True, want_lp=True)
    y_lp.float().pow(2).sum().backward()
    hr = h.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(hr + u.detach().float(), (768,), w.detach(), b.detach(), 1e-5)
    yr.bfloat16().float().pow(2).sum().backward()
    torch.testing.assert_close(h.grad, hr.grad, rtol=2e-2, atol=2e-2)
    with torch.no_grad():
        y2, _ = ops.add_dropout_layer_norm(h, u, w, b, 1e-5, 0.0, training=False)
    torch.testing.assert_close(y2, ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("N,d,dt", [(8128, 2304, torch.bfloat16), (1, 8, torch.float32), (333, 768, torch.float32),
                                    (20928, 3072, torch.bfloat16), (31, 2048, torch.bfloat16)])
def test_col_sum(N, d, dt):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(d)
    x = torch.randn(N, d, generator=g).to(DEV).to(dt)
    out = ops.col_sum(x)
    ref = x.double().sum(0).float()
    assert out.dtype == torch.float32
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3 * max(1.0, N ** 0.5 / 10))


@pytest.mark.parametrize("N,d,udt,ydt,p", [(37, 768, torch.bfloat16, torch.bfloat16, 0.1),
                                           (2048, 768, torch.bfloat16, torch.float32, 0.1),
                                           (5, 64, torch.float32, torch.float32, 0.0),
                                           (100, 1024, torch.float32, torch.float32, 0.2)])
def test_add_dropout_rms_norm_matches_unfused_pair(N, d, udt, ydt, p):
    """fused (hidden + dropout(update), T5LayerNorm(sum)) == residual_dropout_add + rms_norm_residual with the
    same Philox mask; both outputs carry gradients (residual path and norm path)."""
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(N * d)
    h0 = torch.randn(N, d, generator=g).to(DEV)
    u0 = (torch.randn(N, d, generator=g) * 2).to(DEV).to(udt)
    w0 = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV)
    g_res = torch.randn(N, d, generator=g).to(DEV)
    g_y = torch.randn(N, d, generator=g).to(DEV).to(ydt)
    outs = []
    for fused in (True, False):
        h, u, w = (t.clone().requires_grad_(True) for t in (h0, u0, w0))
        ops.manual_seed(21)
        if fused:
            ho, y = ops.add_dropout_rms_norm(h, u, w, 1e-6, p, True, ydt)
        else:
            ho = ops.residual_dropout_add(h, u, p, training=True)
            y, ho = ops.rms_norm_residual(ho, w, 1e-6, ydt)
        assert y.dtype == ydt and ho.dtype == torch.float32
        torch.autograd.backward([ho, y], [g_res, g_y])
        outs.append((ho.detach(), y.detach().float(), h.grad, u.grad.float(), w.grad))
    lp = udt == torch.bfloat16
    tols = {"hidden_out": 1e-6, "y": 2e-2 if ydt == torch.bfloat16 else 2e-5, "d_hidden": 2e-4,
            "d_update": 2e-2 if lp else 2e-4, "d_weight": 2e-3}
    for (n, tol), a, r in zip(tols.items(), outs[0], outs[1]):
        scale = max(1.0, float(r.abs().max())) if n == "d_weight" else 1.0
        torch.testing.assert_close(a, r, rtol=tol, atol=tol * scale, msg=lambda m, n=n: f"{n}: {m}")
    # residual-only consumer (normed output unused) still back-propagates through the dropout mask
    h, u, w = (t.clone().requires_grad_(True) for t in (h0, u0, w0))
    ops.manual_seed(21)
    ho, _ = ops.add_dropout_rms_norm(h, u, w, 1e-6, p, True, ydt)
    ho.backward(g_res)
    torch.testing.assert_close(h.grad, g_res)
    kept = (outs[0][0] - h0) != 0
    torch.testing.assert_close(u.grad.float()[kept], (g_res / (1 - p))[kept], rtol=2e-2, atol=2e-2)
