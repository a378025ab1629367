"""The rest of the model family on CPU: checkpoint layout against the reference's own constructed modules, the
oracle restatements against the reference's recorded outputs, and the host-side beam selection rule.
Fixtures: tests/golden/model_{customizedlatr,customizedprestu,prestu,phonemeprestu}_tiny.npz and
sal_family_layouts.npz, written by oracle/make_golden_variants.py from the real reference classes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_model

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VOCAB = (21, 33, 7)


def _build(name):
    import phoneme_vqa_b200.models as M
    cfg = ref_model.tiny_config()
    if name == "CustomizedLaTr":
        return M.CustomizedLaTr(cfg, tgt_vocab_size=50)
    if name == "CustomizedPreSTU":
        return M.CustomizedPreSTU(cfg, 50)
    if name == "PreSTU":
        return M.PreSTU(cfg)
    return M.PhonemePreSTU(cfg, *VOCAB)


@pytest.mark.parametrize("name", ["CustomizedLaTr", "CustomizedPreSTU", "PreSTU", "PhonemePreSTU"])
def test_state_dict_layout_and_frozen_set_match_reference(name):
    g = np.load(os.path.join(GOLD, f"model_{name.lower()}_tiny.npz"))
    model = _build(name)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state_dict_keys"])
    assert [json.dumps(list(v.shape)) for v in sd.values()] == list(g["state_dict_shapes"])
    assert sorted(k for k, p in model.named_parameters() if not p.requires_grad) == list(g["frozen"])
    # the trainable set is the set of tensors the reference produced gradients for
    trainable = sorted(k for k, p in model.named_parameters() if p.requires_grad)
    unused = {k for k in trainable if k not in set(g["grad_keys"])}
    # parameters the reference registers but never reaches in forward (ViT pooler; tied tables appear once)
    assert all(k.startswith("vit.pooler") or "relative_attention_bias" in k for k in unused), unused
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)


@pytest.mark.parametrize("name,args", [("SaL", ()), ("CustomizedSaL", (50,)), ("PhonemeSaL", (253,))])
def test_sal_family_layouts_match_reference_modules(name, args):
    import phoneme_vqa_b200.models as M
    g = np.load(os.path.join(GOLD, "sal_family_layouts.npz"))
    oracle = getattr(ref_model, name)(ref_model.sal_config(), *args)     # HF mutates the config: one per model
    model = getattr(M, name)(ref_model.sal_config(), *args)
    for m in (oracle, model):
        sd = m.state_dict()
        assert list(sd.keys()) == list(g[name + "_keys"])
        assert [json.dumps(list(v.shape)) for v in sd.values()] == list(g[name + "_shapes"])
    model.load_state_dict(oracle.state_dict(), strict=True)
    oracle.load_state_dict(model.state_dict(), strict=True)
    if name == "SaL":      # tied tables stay tied through a load
        assert model.backbone.encoder.embed_tokens.weight is model.backbone.shared.weight
        assert model.backbone.decoder.embed_tokens.weight is model.backbone.shared.weight


@pytest.mark.parametrize("num_beam", [2, 3])
def test_beam_selection_rule_matches_reference_ids(num_beam):
    """`reference_beam_select` replayed on the scores the reference's beam routine started from."""
    from phoneme_vqa_b200.models import reference_beam_select
    g = np.load(os.path.join(GOLD, "model_customizedlatr_tiny.npz"))
    prob = torch.from_numpy(g["beam_prob"])
    ys = torch.ones(prob.shape[0], 1, dtype=torch.long)
    got = reference_beam_select(prob, ys, end_symbol=2, max_len=5, num_beam=num_beam)
    assert np.array_equal(got.numpy(), g[f"beam{num_beam}_ids"])


def test_beam_selection_rule_edge_cases():
    """eos as the arg-max token: the shared mask zeroes, every beam stops after one more token; negative scores give
    NaN log-probabilities and torch.argmax then selects the first NaN beam (what the reference returns)."""
    from phoneme_vqa_b200.models import reference_beam_select
    prob = torch.tensor([[0.1, 0.2, 5.0, 3.0], [4.0, 0.2, 3.0, 0.5]])
    ys = torch.ones(2, 1, dtype=torch.long)
    got = reference_beam_select(prob, ys, end_symbol=2, max_len=6, num_beam=2)
    # row 0 hits eos at once (its mask entry is 0 from then on), row 1 never does, so the loop runs max_len - 1 times
    assert got.tolist() == [[1, 2, 2, 2, 2, 2, 2], [1, 0, 0, 0, 0, 0, 0]]
    # every row's arg-max is eos: both beams stop after one extra token
    both = reference_beam_select(torch.tensor([[0.1, 0.2, 5.0, 3.0]]), ys[:1], end_symbol=2, max_len=6, num_beam=2)
    assert both.tolist() == [[1, 2, 2]]
    neg = torch.tensor([[-1.0, -2.0, -0.5]])
    got = reference_beam_select(neg, torch.ones(1, 1, dtype=torch.long), end_symbol=9, max_len=3, num_beam=2)
    assert got.tolist() == [[1, 2, 2, 2]]


def test_token_embedding_scales_by_sqrt_d():
    from phoneme_vqa_b200.modules import TokenEmbedding
    te = TokenEmbedding(11, 16)
    ids = torch.tensor([[1, 5, 10]], dtype=torch.int32)
    assert torch.equal(te(ids), te.embedding.weight[ids.long()] * 4.0)
    assert list(te.state_dict().keys()) == ["embedding.weight"]
