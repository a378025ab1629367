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


# ---------------------------------------------------------------------------------------------------------------
# SaL family: the oracle restatement against outputs of the REAL reference classes (tests/golden/model_{sal,
# customizedsal,phonemesal}_tiny.npz; the reference runs through the two-line call-convention adapter documented in
# oracle/make_golden_variants.py)
# ---------------------------------------------------------------------------------------------------------------
def _sal_oracle_case(name):
    cfg = ref_model.sal_config()
    if name == "PhonemeSaL":
        model, batch = ref_model.PhonemeSaL(ref_model.sal_config(), 253), ref_model.sal_batch(3, cfg)
        fwd = lambda m: m(batch)[0]                    # noqa: E731
        loss = lambda m: m(batch)[1]                   # noqa: E731
    elif name == "CustomizedSaL":
        model, batch = ref_model.CustomizedSaL(ref_model.sal_config(), 50), ref_model.customized_sal_batch(3, cfg)
        fwd = lambda m: m(batch)                       # noqa: E731
        loss = lambda m: ref_model.sal_t5_loss(m, batch)   # noqa: E731
    else:
        model, batch = ref_model.SaL(ref_model.sal_config()), ref_model.sal_t5_batch(3, cfg)
        fwd = lambda m: m(batch)                       # noqa: E731
        loss = lambda m: ref_model.sal_t5_loss(m, batch)   # noqa: E731
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    return model, batch, fwd, loss


@pytest.mark.parametrize("name", ["PhonemeSaL", "CustomizedSaL", "SaL"])
def test_sal_family_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"model_{name.lower()}_tiny.npz"))
    model, batch, fwd, loss_fn = _sal_oracle_case(name)
    assert list(model.state_dict().keys()) == list(g["state_dict_keys"])
    model.eval()
    with torch.no_grad():
        np.testing.assert_allclose(fwd(model).numpy(), g["logits"], rtol=1e-5, atol=1e-6)
        if name == "PhonemeSaL":
            assert np.array_equal(model.generate(batch, 1, 2, max_len=5).numpy(), g["greedy_ids"])
        elif name == "CustomizedSaL":
            assert np.array_equal(model.greedy_generate(batch, 1, 2, 6).numpy(), g["greedy_ids"])
        else:
            assert np.array_equal(model.generate(batch, max_length=6).numpy(), g["generate_ids"])
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0
    loss = loss_fn(model)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-6)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads) == list(g["grad_keys"])
    norms = np.array([grads[k].double().norm().item() for k in sorted(grads)])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)


def test_customized_sal_beam_ids_follow_the_reference_rule():
    """the reference's CustomizedSaL.beam_generate output, replayed from the oracle's first-step scores"""
    from phoneme_vqa_b200.models import reference_beam_select
    g = np.load(os.path.join(GOLD, "model_customizedsal_tiny.npz"))
    model, batch, _, _ = _sal_oracle_case("CustomizedSaL")
    model.eval()
    with torch.no_grad():
        enc, mask = model._encode(batch)
        ys = torch.ones(3, 1, dtype=torch.long)
        prob = model.lm_head(model.decode(ys, enc, mask)[:, -1])
    assert np.array_equal(reference_beam_select(prob, ys, 2, 4, 2).numpy(), g["beam2_ids"])


# ---------------------------------------------------------------------------------------------------------------
# factorised beam search: the batched, cache-reordering implementation against a naive
# per-sample search over the FULL onset x rhyme x tone product
# ---------------------------------------------------------------------------------------------------------------
class _ToyDecoder:
    """stateful stand-in for the cached decoder step: the log-probs of a row depend on the whole history of that
    row (a running hash), so a wrong cache reorder changes the result"""

    def __init__(self, bz, K, vocab, seed):
        g = torch.Generator().manual_seed(seed)
        self.V = vocab
        self.W = [torch.randn(64, v, generator=g) * 2.0 for v in vocab]
        self.sample_bias = torch.randn(bz, 64, generator=g)
        self.K = K
        self.state = None

    def _feat(self, state):
        idx = torch.arange(64)[None, :]
        return torch.sin(state[:, None].double() * (idx + 1) * 0.37).float()

    def logp(self, state, sample):
        f = self._feat(state) + self.sample_bias[sample]
        return [torch.log_softmax(f @ w, dim=-1) for w in self.W]

    @staticmethod
    def advance(state, tok):
        return (state * 31 + tok[:, 0] * 7 + tok[:, 1] * 3 + tok[:, 2] + 1) % 1000003

    def step(self, tok, t, src):
        n = tok.shape[0]
        sample = torch.arange(n) // self.K
        if self.state is None:
            self.state = torch.zeros(n, dtype=torch.long)
        if src is not None:
            self.state = self.state[src]
        self.state = self.advance(self.state, tok[:, 0])
        return tuple(self.logp(self.state, sample))


def _naive_beam(toy, sample, K, start, end, max_len):
    beams = [(0.0, [[start, 0, 0]], False, toy.advance(torch.zeros(1, dtype=torch.long), torch.tensor([[start, 0, 0]])))]
    for _ in range(max_len):
        cands = []
        for score, seq, fin, state in beams:
            if fin:
                cands.append((score, seq + [[end, 0, 0]], True, state))
                continue
            on, rh, to = toy.logp(state, torch.tensor([sample]))
            for o in range(toy.V[0]):
                for r in range(toy.V[1]):
                    for c in range(toy.V[2]):
                        s = score + float(on[0, o]) + float(rh[0, r]) + float(to[0, c])
                        tok = [o, r, c]
                        cands.append((s, seq + [tok], o == end, toy.advance(state, torch.tensor([tok]))))
        cands.sort(key=lambda x: -x[0])
        beams = cands[:K]
        if all(b[2] for b in beams):
            break
    return beams[0][1], beams[0][0]


@pytest.mark.parametrize("K", [1, 2, 3, 5])
def test_factorised_beam_search_equals_full_product_search(K):
    from phoneme_vqa_b200.models import phoneme_beam_search
    bz, vocab, start, end, max_len = 3, (6, 5, 4), 0, 1, 7
    toy = _ToyDecoder(bz, K, vocab, seed=10 + K)
    got = phoneme_beam_search(toy.step, bz, K, start, end, max_len, torch.device("cpu"))
    assert got.dtype == torch.long and got.shape[0] == bz and got.shape[2] == 3
    for b in range(bz):
        want, _ = _naive_beam(toy, b, K, start, end, max_len)
        assert got[b, : len(want)].tolist() == want, (b, got[b].tolist(), want)


def test_beam_search_with_one_beam_is_greedy():
    from phoneme_vqa_b200.models import phoneme_beam_search
    bz, vocab = 4, (7, 6, 3)
    toy = _ToyDecoder(bz, 1, vocab, seed=3)
    got = phoneme_beam_search(toy.step, bz, 1, 0, 1, 6, torch.device("cpu"))
    state = toy.advance(torch.zeros(bz, dtype=torch.long), torch.tensor([[0, 0, 0]] * bz))
    done = torch.zeros(bz, dtype=torch.bool)
    for t in range(1, got.shape[1]):
        on, rh, to = toy.logp(state, torch.arange(bz))
        tok = torch.stack([on.argmax(-1), rh.argmax(-1), to.argmax(-1)], dim=-1)
        tok = torch.where(done[:, None], torch.tensor([[1, 0, 0]]), tok)
        assert torch.equal(got[:, t], tok)
        done = done | (tok[:, 0] == 1)
        state = toy.advance(state, tok)


# ---------------------------------------------------------------------------------------------------------------
# synthetic batches of the sibling bench workloads (bench.py --workload phonoprestu / phonosal)
# ---------------------------------------------------------------------------------------------------------------
def test_sibling_workload_batches_have_the_dataset_contract():
    from phoneme_vqa_b200 import synthetic
    b = synthetic.phoneme_prestu_batch(3, 500, T=9, L_in=14, image=32)
    assert set(b) == {"pixel_values", "input_ids", "src_attention_mask", "label_ids", "label_attention_mask"}
    assert b["input_ids"].shape == (3, 14) and b["input_ids"].dtype == torch.int64
    assert b["label_ids"].shape == (3, 10, 3) and b["pixel_values"].shape == (3, 3, 32, 32)
    assert torch.equal(b["label_ids"][:, 0], torch.tensor([[synthetic.BOS_ID, 0, 0]] * 3))
    s = synthetic.phoneme_sal_batch(3, 500, T=9, L_q=16, L_ocr=32, L_obj=16, ocr_hidden=24, obj_hidden=40)
    assert tuple(s) == synthetic.SAL_FIELDS or set(s) == set(synthetic.SAL_FIELDS)
    assert s["label_attention_mask"].dtype == torch.bool and s["ocr_coordinates"].shape == (3, 32, 4)
    assert float(s["ocr_coordinates"].max()) < 1.0 and float(s["ocr_coordinates"].min()) >= 0.0
    assert torch.equal(s["label_ids"][:, 1:], s["shifted_right_label_ids"][:, :-1])        # shifted by one
    assert (s["label_ids"][:, 0] == 1).all() and ((s["shifted_right_label_ids"] == 2).sum(1) == 1).all()
    # every sequence field ends with </s> (id 1) then pads, and the float masks cover exactly the non-pad prefix
    for ids, mask in ((s["input_ids"], s["src_attention_mask"]), (s["tokenized_ocr"], s["ocr_attention_mask"])):
        assert torch.equal(mask.bool(), ids != 0)
