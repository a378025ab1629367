"""K4 parity: fused phoneme head + 3xCE (CUDA, through the C-ABI) vs the CPU oracle."""
import pytest
import torch

from oracle import ref_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(N, d, V=(84, 187, 7), seed=0, ignore=2, frac_ignored=0.6):
    g = torch.Generator().manual_seed(seed)
    rt = d // 3
    on = d - 2 * rt
    h = torch.randn(N, d, generator=g)
    tg = torch.stack([torch.randint(0, V[k], (N,), generator=g) for k in range(3)], dim=-1)
    ign = torch.rand(N, generator=g) < frac_ignored
    tg[ign] = ignore
    Ws = [torch.randn(V[0], on, generator=g) * 0.1, torch.randn(V[1], rt, generator=g) * 0.1,
          torch.randn(V[2], rt, generator=g) * 0.1]
    bs = [torch.randn(V[k], generator=g) * 0.1 for k in range(3)]
    return h, tg, Ws, bs


@pytest.mark.parametrize("N,d", [(37, 192), (8128, 768), (1, 48), (300, 1032)])
def test_phoneme_head_ce_fp32(N, d):
    from phoneme_vqa_b200 import ops
    h, tg, Ws, bs = _inputs(N, d, seed=N)
    leaves = [h] + Ws + bs
    for t in leaves:
        t.requires_grad_(True)
    ref, _ = ref_ops.phoneme_head_ce(h, tg, Ws[0], bs[0], Ws[1], bs[1], Ws[2], bs[2], ignore_index=2)
    ref.backward()
    c = [t.detach().to(DEV).requires_grad_(True) for t in leaves]
    loss = ops.phoneme_head_ce(c[0], tg.to(DEV), c[1], c[4], c[2], c[5], c[3], c[6], 2)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-5, atol=1e-6, equal_nan=True)
    for a, b in zip(c, leaves):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-3, atol=1e-6, equal_nan=True)


def test_phoneme_head_ce_bf16():
    from phoneme_vqa_b200 import ops
    h, tg, Ws, bs = _inputs(2000, 768, seed=3)
    ref, _ = ref_ops.phoneme_head_ce(h.bfloat16().float(), tg, *[x for pair in zip([w.bfloat16().float() for w in Ws],
                                                                                   [b.bfloat16().float() for b in bs])
                                                                  for x in pair], ignore_index=2)
    loss = ops.phoneme_head_ce(h.to(DEV).bfloat16(), tg.to(DEV), Ws[0].to(DEV), bs[0].to(DEV), Ws[1].to(DEV),
                               bs[1].to(DEV), Ws[2].to(DEV), bs[2].to(DEV), 2)
    assert abs(loss.item() - ref.item()) <= 1e-3 * abs(ref.item())


def test_phoneme_head_ce_strided_targets_and_all_ignored_head():
    from phoneme_vqa_b200 import ops
    h, tg, Ws, bs = _inputs(64, 192, seed=9, frac_ignored=0.0)
    labels = torch.zeros(4, 17, 3, dtype=torch.long)
    labels[:, 1:] = tg.view(4, 16, 3)
    tview = labels[:, 1:].reshape(-1, 3)                     # executor's labels[:, 1:, k]
    ref, _ = ref_ops.phoneme_head_ce(h, tview, Ws[0], bs[0], Ws[1], bs[1], Ws[2], bs[2], ignore_index=-100)
    loss = ops.phoneme_head_ce(h.to(DEV), tview.to(DEV), Ws[0].to(DEV), bs[0].to(DEV), Ws[1].to(DEV), bs[1].to(DEV),
                               Ws[2].to(DEV), bs[2].to(DEV), -100)
    torch.testing.assert_close(loss.cpu(), ref, rtol=1e-5, atol=1e-6)


# ------------------------------- K4 large-vocabulary variant (LaTr) ------------------------------------
@pytest.mark.parametrize("N,d,V,dtype", [(50, 64, 120, torch.float32), (300, 192, 1000, torch.float32),
                                         (2500, 768, 36096, torch.bfloat16), (7, 32, 37, torch.float32)])
def test_vocab_head_ce_chunked(N, d, V, dtype):
    from phoneme_vqa_b200 import ops
    g = torch.Generator().manual_seed(N)
    h = torch.randn(N, d, generator=g)
    W = torch.randn(V, d, generator=g) * 0.05
    tg = torch.randint(0, V, (N,), generator=g)
    tg[torch.rand(N, generator=g) < 0.5] = 0            # pad id 0 ignored
    hr = h.to(dtype).float().clone().requires_grad_(True)
    Wr = (W.to(dtype).float() if dtype == torch.bfloat16 else W.clone()).requires_grad_(True)
    ref, _ = ref_ops.vocab_head_ce(hr, Wr, tg, ignore_index=0)
    ref.backward()
    hd = h.detach().to(DEV).to(dtype).requires_grad_(True)
    Wd = W.detach().to(DEV).requires_grad_(True)
    loss = ops.vocab_head_ce(hd, Wd, tg.to(DEV), 0, w_lp=(Wd.detach().to(dtype) if dtype != torch.float32 else None),
                             chunk_rows=1024 if N > 1024 else 64)
    loss.backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=tol, atol=tol)
    gtol = 2e-2 if dtype == torch.float32 else 5e-2    # dlogits are stored in bf16
    assert (hd.grad.float().cpu() - hr.grad).norm() / hr.grad.norm() <= gtol
    assert (Wd.grad.float().cpu() - Wr.grad).norm() / Wr.grad.norm() <= gtol
