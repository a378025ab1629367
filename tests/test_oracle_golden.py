"""Pin the oracle (CPU restatement) to the reference's own outputs (tests/golden/*, written by
oracle/make_golden.py from /root/reference).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import ref_model, ref_ops

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _disable_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0


def test_ops_golden():
    g = np.load(os.path.join(GOLD, "ops_small.npz"))
    cfg = ref_model.tiny_config(d_model=48, d_kv=16, num_heads=3)
    sm = ref_model.SpatialModule(cfg)
    sm.load_state_dict(ref_model.deterministic_state_dict(sm, scale=1.0))
    out = sm(torch.from_numpy(g["spatial_coords"]))
    assert np.array_equal(out.detach().numpy(), g["spatial_out"])
    assert np.array_equal(ref_ops.sinusoidal_table(48, 64).numpy(), g["pe_table"])
    x = torch.from_numpy(g["pe_in"])
    assert np.array_equal(ref_ops.positional_encoding(x, torch.from_numpy(g["pe_table"])).numpy(), g["pe_out"])


def test_model_golden_forward_loss_grads_greedy():
    g = np.load(os.path.join(GOLD, "model_phonemelatr_tiny.npz"))
    cfg = ref_model.tiny_config()
    vocab = (21, 33, 7)
    model = ref_model.PhonemeLaTr(cfg, *vocab)
    assert list(model.state_dict().keys()) == list(g["state_dict_keys"])
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    batch = ref_model.synthetic_batch(3, cfg, T=9, L_ocr=12, L_q=6, V_sub=vocab, seed=7, image=32)
    model.eval()
    labels = batch["label_ids"]
    on, rh, to = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"],
                       input_ids=batch["input_ids"], labels=labels[:, :-1],
                       src_attention_mask=batch["src_attention_mask"],
                       label_attention_mask=batch["label_attention_mask"][:, :-1],
                       ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    for got, key in ((on, "onset_logits"), (rh, "rhyme_logits"), (to, "tone_logits")):
        np.testing.assert_allclose(got.detach().numpy(), g[key], rtol=1e-5, atol=1e-6)
    model.train()
    _disable_dropout(model)
    loss = ref_model.phoneme_latr_loss(model, batch, 2)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-6)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads.keys()) == list(g["grad_keys"])
    norms = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)
    for key in g.files:
        if key.startswith("grad::"):
            np.testing.assert_allclose(grads[key[6:]].numpy(), g[key], rtol=1e-4, atol=1e-7)
    model.eval()
    ys = model.greedy_generate(batch["pixel_values"], batch["coordinates"], batch["input_ids"],
                               batch["src_attention_mask"], batch["ocr_attention_mask"], batch["tokenized_ocr"],
                               start_symbol=3, end_symbol=4, max_len=6)
    assert np.array_equal(ys.numpy(), g["greedy_ids"])          # index tensors: bit-exact


def test_product_model_state_dict_layout_matches_reference():
    """checkpoint contract: same keys, same order, same shapes as the reference model."""
    import phoneme_vqa_b200.models as M
    g = np.load(os.path.join(GOLD, "model_phonemelatr_tiny.npz"))
    cfg = ref_model.tiny_config()
    model = M.PhonemeLaTr(cfg, 21, 33, 7)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state_dict_keys"])
    assert [json.dumps(list(v.shape)) for v in sd.values()] == list(g["state_dict_shapes"])
    oracle = ref_model.PhonemeLaTr(cfg, 21, 33, 7)
    model.load_state_dict(oracle.state_dict(), strict=True)       # reference -> product
    oracle.load_state_dict(model.state_dict(), strict=True)       # product -> reference
    frozen = {k for k, p in model.named_parameters() if not p.requires_grad}
    assert frozen == {k for k, p in oracle.named_parameters() if not p.requires_grad}


def test_latr_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "model_latr_tiny.npz"))
    cfg = ref_model.tiny_config()
    model = ref_model.LaTr(cfg)
    assert list(model.state_dict().keys()) == list(g["state_dict_keys"])
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    batch = ref_model.latr_batch(3, cfg)
    model.eval()
    labels = batch["label_ids"]
    logits = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"], input_ids=batch["input_ids"],
                   labels=labels[:, :-1], src_attention_mask=batch["src_attention_mask"],
                   label_attention_mask=batch["label_attention_mask"][:, :-1],
                   ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    np.testing.assert_allclose(logits.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    model.train()
    _disable_dropout(model)
    loss = ref_model.latr_loss(model, batch)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-6)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads.keys()) == list(g["grad_keys"])
    norms = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)
    model.eval()
    ys = model.generate(batch["pixel_values"], batch["coordinates"], batch["input_ids"], batch["src_attention_mask"],
                        batch["ocr_attention_mask"], batch["tokenized_ocr"], max_length=8)
    assert np.array_equal(ys.numpy(), g["generate_ids"])


def test_product_latr_state_dict_layout_matches_reference():
    import phoneme_vqa_b200.models as M
    g = np.load(os.path.join(GOLD, "model_latr_tiny.npz"))
    cfg = ref_model.tiny_config()
    model = M.LaTr(cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state_dict_keys"])
    assert [json.dumps(list(v.shape)) for v in sd.values()] == list(g["state_dict_shapes"])
    oracle = ref_model.LaTr(cfg)
    model.load_state_dict(oracle.state_dict(), strict=True)
    oracle.load_state_dict(model.state_dict(), strict=True)
    # tied weights stay tied after loading
    assert model.backbone.lm_head.weight is model.backbone.shared.weight
    assert model.backbone.decoder.embed_tokens.weight is model.backbone.shared.weight


def test_sal_bias_oracle_and_product_match_reference_modules():
    """the reference's RelativePositionBiasAggregated output (real modules, CPU) vs the oracle restatement and
    vs the product's on-device formulation (relative vector + uint8 SCP bucket map), bit-exact."""
    g = np.load(os.path.join(GOLD, "sal_bias.npz"))
    S, q0, L = int(g["S"]), int(g["q0"]), int(g["L"])
    coords = torch.from_numpy(g["coords"])
    rel_t, scp_t = torch.from_numpy(g["rel_table"]), torch.from_numpy(g["scp_table"])
    ora = ref_model.sal_position_bias(rel_t, scp_t, S, coords, q0, L)
    assert np.array_equal(ora.numpy(), g["bias"])
    import phoneme_vqa_b200.modules as PM
    agg = PM.RelativePositionBiasAggregated(PM.RelativePositionBias1D(int(g["H"])), PM.SCPRelativePositionBias(int(g["H"])))
    assert list(agg.state_dict().keys()) == list(g["state_dict_keys"])
    agg.Relative1D.relative_attention_bias.weight.data.copy_(rel_t)
    agg.SCP.relative_attention_bias.weight.data.copy_(scp_t)
    dense = agg.dense(S, coords, q0, L)
    assert np.array_equal(dense.detach().numpy(), g["bias"])
    rel, (bk, tab, qq) = agg(S, coords, q0, L)
    assert bk.dtype == torch.uint8 and bk.shape == (2, L, L) and rel.shape == (int(g["H"]), 2 * S - 1) and qq == q0


def test_product_sal_state_dict_layout_matches_oracle():
    import phoneme_vqa_b200.models as M
    cfg = ref_model.sal_config()
    oracle = ref_model.PhonemeSaL(cfg, 253)
    model = M.PhonemeSaL(cfg, 253)
    assert list(model.state_dict().keys()) == list(oracle.state_dict().keys())
    assert [tuple(v.shape) for v in model.state_dict().values()] == [tuple(v.shape) for v in oracle.state_dict().values()]
    model.load_state_dict(oracle.state_dict(), strict=True)
    oracle.load_state_dict(model.state_dict(), strict=True)
