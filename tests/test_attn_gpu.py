"""K2 / K3 parity: tcgen05 flash attention (through the C-ABI) vs the CPU oracle restatements of
HF T5Attention's core and nn.MultiheadAttention's core (oracle/ref_ops.py)."""
import math

import pytest
import torch

from oracle import ref_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _qkv(B, Sq, Sk, H, D=64, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, Sq, H, D, generator=g) * scale).bfloat16()
    k = (torch.randn(B, Sk, H, D, generator=g) * scale).bfloat16()
    v = torch.randn(B, Sk, H, D, generator=g).bfloat16()
    return q, k, v


def _t(x):  # (B,S,H,D) -> (B,H,S,D) float
    return x.float().transpose(1, 2)


def _rel_dense(rel, Sq, Sk):
    i = torch.arange(Sq)[:, None]
    j = torch.arange(Sk)[None, :]
    return rel[:, (j - i + Sq - 1)]


T5_CASES = [(1, 128, 128, 1), (2, 327, 327, 3), (1, 100, 100, 2), (2, 464, 464, 2), (1, 707, 707, 1), (3, 1, 5, 2),
            (1, 129, 257, 1)]


@pytest.mark.parametrize("B,Sq,Sk,H", T5_CASES)
def test_t5_attention_fwd(B, Sq, Sk, H):
    from phoneme_vqa_b200 import ops
    q, k, v = _qkv(B, Sq, Sk, H, seed=Sq, scale=0.35)
    g = torch.Generator().manual_seed(1)
    rel = torch.randn(H, Sq + Sk - 1, generator=g)
    valid = torch.rand(B, Sk, generator=g) > 0.2
    valid[:, 0] = True
    ref = ref_ops.t5_attention_core(_t(q), _t(k), _t(v), _rel_dense(rel, Sq, Sk)[None], key_valid=valid)
    key_add = torch.where(valid, 0.0, float("-inf"))
    o, lse = ops.attention_fwd_raw(q.to(DEV), k.to(DEV), v.to(DEV), 1.0, rel.to(DEV), key_add.to(DEV))
    got = o.float().cpu().transpose(1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2e-2, err                       # bf16 P and bf16 output
    assert (got - ref).norm() / ref.norm() <= 1e-2
    # lse against fp64 scores
    s = torch.einsum("bhid,bhjd->bhij", _t(q).double(), _t(k).double()) + _rel_dense(rel, Sq, Sk)[None].double()
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    torch.testing.assert_close(lse.cpu().double(), torch.logsumexp(s, -1), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,T,S,H", [(2, 127, 327, 3), (1, 40, 64, 2), (2, 9, 17, 1)])
def test_mha_cross_attention_fwd_float_masks(B, T, S, H):
    from phoneme_vqa_b200 import ops
    q, k, v = _qkv(B, T, S, H, seed=T)
    g = torch.Generator().manual_seed(2)
    key_add = (torch.rand(B, S, generator=g) > 0.3).float()       # reference passes 1.0 = valid, ADDED (D14)
    ref = ref_ops.mha_attention_core(_t(q), _t(k), _t(v), causal=False, key_add=key_add)
    o, _ = ops.attention_fwd_raw(q.to(DEV), k.to(DEV), v.to(DEV), 1.0 / math.sqrt(64), None, key_add.to(DEV))
    got = o.float().cpu().transpose(1, 2)
    assert (got - ref).abs().max().item() <= 2e-2
    assert (got - ref).norm() / ref.norm() <= 1e-2


@pytest.mark.parametrize("B,T,H", [(2, 127, 3), (1, 128, 1), (1, 300, 2), (2, 1, 1)])
def test_mha_causal_self_attention_fwd_packed_qkv(B, T, H):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, T, 3, H, 64, generator=g).bfloat16()
    key_add = (torch.rand(B, T, generator=g) > 0.7).float()       # 1.0 = pad, ADDED
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ref = ref_ops.mha_attention_core(_t(q), _t(k), _t(v), causal=True, key_add=key_add)
    qkv_d = qkv.to(DEV)
    o, _ = ops.attention_fwd_raw(qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2], 1.0 / math.sqrt(64), None,
                                 key_add.to(DEV), causal=True)
    got = o.float().cpu().transpose(1, 2)
    assert (got - ref).abs().max().item() <= 2e-2
    assert (got - ref).norm() / ref.norm() <= 1e-2


# ------------------------------- backward -------------------------------------------
def _ref_grads(q, k, v, scale, rel, key_add, causal, go):
    """fp32 autograd through the oracle-equivalent math (scores exactly as the kernels define them)."""
    qf, kf, vf = [t.float().transpose(1, 2).clone().requires_grad_(True) for t in (q, k, v)]
    relp = None if rel is None else rel.clone().requires_grad_(True)
    Sq, Sk = qf.shape[2], kf.shape[2]
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if relp is not None:
        s = s + _rel_dense(relp, Sq, Sk)[None]
    if key_add is not None:
        s = s + key_add[:, None, None, :]
    if causal:
        s = s + torch.full((Sq, Sk), float("-inf")).triu(1)
    o = torch.matmul(torch.softmax(s, -1), vf)
    o.backward(go.float().transpose(1, 2))
    tr = lambda t: t.grad.transpose(1, 2)  # noqa: E731
    return o.detach().transpose(1, 2), tr(qf), tr(kf), tr(vf), (None if relp is None else relp.grad)


def _close(a, b, tol=2e-2):
    a = a.float().cpu()
    if b.norm() < 1e-6:                      # e.g. a single key: dS is identically zero
        assert a.norm() < 1e-3, float(a.norm())
        return
    err = (a - b).norm() / b.norm()
    assert err <= tol, float(err)


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (2, 327, 3), (1, 100, 2), (1, 464, 2), (2, 5, 1)])
def test_t5_self_attention_bwd_packed(B, S, H):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B, S, 3, H, 64, generator=g) * 0.5).bfloat16()
    rel = torch.randn(H, 2 * S - 1, generator=g)
    valid = torch.rand(B, S, generator=g) > 0.2
    valid[:, 0] = True
    key_add = torch.where(valid, 0.0, float("-inf"))
    go = torch.randn(B, S, H, 64, generator=g).bfloat16()
    _, dq, dk, dv, drel = _ref_grads(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 1.0, rel, key_add, False, go)
    qd = qkv.to(DEV).requires_grad_(True)
    reld = rel.to(DEV).requires_grad_(True)
    o = ops.attention_self(qd, 1.0, rel_bias=reld, key_add=key_add.to(DEV))
    o.backward(go.to(DEV))
    _close(qd.grad[:, :, 0], dq)
    _close(qd.grad[:, :, 1], dk)
    _close(qd.grad[:, :, 2], dv)
    _close(reld.grad, drel)


@pytest.mark.parametrize("B,T,H", [(2, 127, 3), (1, 300, 2), (2, 1, 1)])
def test_mha_causal_self_attention_bwd(B, T, H):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, T, 3, H, 64, generator=g).bfloat16()
    key_add = (torch.rand(B, T, generator=g) > 0.7).float()
    go = torch.randn(B, T, H, 64, generator=g).bfloat16()
    sc = 1.0 / math.sqrt(64)
    _, dq, dk, dv, _ = _ref_grads(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], sc, None, key_add, True, go)
    qd = qkv.to(DEV).requires_grad_(True)
    o = ops.attention_self(qd, sc, key_add=key_add.to(DEV), causal=True)
    o.backward(go.to(DEV))
    _close(qd.grad[:, :, 0], dq)
    _close(qd.grad[:, :, 1], dk)
    _close(qd.grad[:, :, 2], dv)


@pytest.mark.parametrize("B,T,S,H", [(2, 127, 327, 3), (1, 40, 64, 2)])
def test_mha_cross_attention_bwd(B, T, S, H):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(T + S)
    q = torch.randn(B, T, H, 64, generator=g).bfloat16()
    kv = torch.randn(B, S, 2, H, 64, generator=g).bfloat16()
    key_add = (torch.rand(B, S, generator=g) > 0.3).float()
    go = torch.randn(B, T, H, 64, generator=g).bfloat16()
    sc = 1.0 / math.sqrt(64)
    _, dq, dk, dv, _ = _ref_grads(q, kv[:, :, 0], kv[:, :, 1], sc, None, key_add, False, go)
    qd = q.to(DEV).requires_grad_(True)
    kvd = kv.to(DEV).requires_grad_(True)
    o = ops.attention_cross(qd, kvd, sc, key_add=key_add.to(DEV))
    o.backward(go.to(DEV))
    _close(qd.grad, dq)
    _close(kvd.grad[:, :, 0], dk)
    _close(kvd.grad[:, :, 1], dv)


# ------------------------------- fp32 parity-mode kernels -----------------------------
@pytest.mark.parametrize("B,S,H,causal", [(2, 37, 2, False), (1, 130, 3, True), (2, 327, 1, False)])
def test_fp32_attention_fwd_bwd_tight(B, S, H, causal):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B, S, 3, H, 64, generator=g) * 0.5
    rel = None if causal else torch.randn(H, 2 * S - 1, generator=g)
    key_add = (torch.rand(B, S, generator=g) > 0.7).float()
    go = torch.randn(B, S, H, 64, generator=g)
    sc = 0.125 if causal else 1.0
    o_ref, dq, dk, dv, drel = _ref_grads(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], sc, rel, key_add, causal, go)
    qd = qkv.to(DEV).requires_grad_(True)
    reld = None if rel is None else rel.to(DEV).requires_grad_(True)
    o = ops.attention_self(qd, sc, rel_bias=reld, key_add=key_add.to(DEV), causal=causal)
    torch.testing.assert_close(o.detach().cpu(), o_ref, rtol=1e-5, atol=1e-5)
    o.backward(go.to(DEV))
    torch.testing.assert_close(qd.grad[:, :, 0].cpu(), dq, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(qd.grad[:, :, 1].cpu(), dk, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(qd.grad[:, :, 2].cpu(), dv, rtol=1e-4, atol=1e-5)
    if rel is not None:
        torch.testing.assert_close(reld.grad.cpu(), drel, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_dropout_mask_is_consistent_between_fwd_and_bwd(dtype):
    """With V = one-hot rows the output IS the dropped probability matrix, so the keep-mask can be read
    back; the backward must then equal autograd through softmax * mask / keep."""
    from phoneme_vqa_b200 import ops
    B, S, H, p = 2, 64, 2, 0.25
    g = torch.Generator().manual_seed(5)
    q = torch.randn(B, S, H, 64, generator=g) * 0.3
    k = torch.randn(B, S, H, 64, generator=g) * 0.3
    v = torch.eye(64)[None, :, None, :].expand(B, S, H, 64).contiguous()
    qkv = torch.stack([q, k, v], dim=2).to(dtype)
    go = torch.randn(B, S, H, 64, generator=g).to(dtype)
    ops.manual_seed(77)
    qd = qkv.to(DEV).requires_grad_(True)
    o = ops.attention_self(qd, 1.0, dropout_p=p)
    pd = o.detach().float().cpu().transpose(1, 2)                 # (B,H,S,S) dropped probabilities
    mask = (pd != 0).float()
    keep = 1.0 - round(p * 256) / 256.0
    assert abs(mask.mean().item() - keep) < 2e-2
    qf, kf = [t.float().transpose(1, 2).requires_grad_(True) for t in (qkv[:, :, 0], qkv[:, :, 1])]
    vf = qkv[:, :, 2].float().transpose(1, 2).requires_grad_(True)
    pr = torch.softmax(torch.matmul(qf, kf.transpose(-1, -2)), -1) * mask / keep
    torch.testing.assert_close(pd, pr.detach(), rtol=2e-2, atol=2e-3)
    torch.matmul(pr, vf).backward(go.float().transpose(1, 2))
    o.backward(go.to(DEV))
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    for idx, ref in ((0, qf), (1, kf), (2, vf)):
        _close(qd.grad[:, :, idx], ref.grad.transpose(1, 2), tol)
    # same seed -> same mask; next call -> different mask
    ops.manual_seed(77)
    o2 = ops.attention_self(qkv.to(DEV), 1.0, dropout_p=p)
    assert torch.equal(o2, o.detach())
    o3 = ops.attention_self(qkv.to(DEV), 1.0, dropout_p=p)
    assert not torch.equal(o3 != 0, o2 != 0)


# ------------------------------- SaL spatial (SCP) bias in-kernel ---------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_self_attention_with_scp_bias_fwd_bwd(dtype):
    from phoneme_vqa_b200 import ops
    B, S, H, q0, L = 2, 208, 2, 16, 128
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(B, S, 3, H, 64, generator=g) * 0.4).to(dtype)
    rel = torch.randn(H, 2 * S - 1, generator=g)
    tab = torch.randn(32, H, generator=g)
    bk = torch.randint(0, 32, (B, L, L), generator=g, dtype=torch.int64).to(torch.uint8)
    go = torch.randn(B, S, H, 64, generator=g).to(dtype)
    # reference math: dense (B,H,S,S) bias, no key mask (SaL passes the bias externally)
    qf, kf, vf = [qkv[:, :, i].float().transpose(1, 2).clone().requires_grad_(True) for i in range(3)]
    relp, tabp = rel.clone().requires_grad_(True), tab.clone().requires_grad_(True)
    dense = _rel_dense(relp, S, S)[None].repeat(B, 1, 1, 1)
    scp = tabp[bk.long()].permute(0, 3, 1, 2)
    dense = torch.cat([dense[:, :, :q0], torch.cat([dense[:, :, q0:q0 + L, :q0], dense[:, :, q0:q0 + L, q0:q0 + L] + scp,
                                                    dense[:, :, q0:q0 + L, q0 + L:]], dim=3), dense[:, :, q0 + L:]], dim=2)
    o_ref = torch.matmul(torch.softmax(torch.matmul(qf, kf.transpose(-1, -2)) + dense, -1), vf)
    o_ref.backward(go.float().transpose(1, 2))
    qd = qkv.to(DEV).requires_grad_(True)
    reld, tabd = rel.to(DEV).requires_grad_(True), tab.to(DEV).requires_grad_(True)
    o = ops.attention_self(qd, 1.0, rel_bias=reld, scp=(bk.to(DEV), tabd, q0))
    o.backward(go.to(DEV))
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    _close(o.detach().transpose(1, 2), o_ref.detach(), tol)
    for idx, r in ((0, qf), (1, kf), (2, vf)):
        _close(qd.grad[:, :, idx], r.grad.transpose(1, 2), tol if dtype == torch.float32 else 3e-2)
    _close(reld.grad, relp.grad, tol if dtype == torch.float32 else 3e-2)
    _close(tabd.grad, tabp.grad, tol if dtype == torch.float32 else 3e-2)
