"""Text pipeline (tokenisation / phoneme index tensors) — must be BIT-EXACT against the reference
(tests/golden/text_golden.json, written from /root/reference by oracle/make_golden_text.py)."""
import hashlib
import json
import os
import tempfile

import pytest

from oracle import text_cases

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "text_golden.json"), encoding="utf-8"))


def _digest(items):
    h = hashlib.sha256()
    for it in items:
        h.update(json.dumps(it, ensure_ascii=False, sort_keys=True).encode("utf-8"))
        h.update(b"\n")
    return h.hexdigest()


def _plain(r):
    return [list(x) if isinstance(x, tuple) else x for x in r]


@pytest.fixture(scope="module")
def text():
    import phoneme_vqa_b200.text as T
    return T


def test_word_analysis_both_variants_bit_exact(text):
    words = text_cases.all_words()
    assert len(words) == GOLD["n_words"]
    n = len(GOLD["explicit_words"])
    assert words[:n] == GOLD["explicit_words"]
    core = [_plain(text.analyse(w, "core")) for w in words]
    dec = [_plain(text.analyse(w, "decode")) for w in words]
    for w, got, ref in zip(words[:n], core[:n], GOLD["explicit_core"]):
        assert got == ref, (w, got, ref)
    for w, got, ref in zip(words[:n], dec[:n], GOLD["explicit_decode"]):
        assert got == ref, (w, got, ref)
    assert _digest(core) == GOLD["core_sha256"]
    assert _digest(dec) == GOLD["decode_sha256"]
    composed = [text.compose_word(*r[1]) if r[0] else None for r in dec]
    assert composed[:n] == GOLD["explicit_compose"]
    assert _digest(composed) == GOLD["compose_sha256"]


def test_survey_seed_examples(text):
    assert text.analyse("quyển", "core") == (True, ("q", "uyên", "<?>"))
    assert text.analyse("giếng", "core") == (True, ("gi", "iêng", "</>"))
    assert text.analyse("gì", "core") == (True, ("gi", "i", "<`>"))
    assert text.analyse("coca", "core") == (False, ("", "", "coca"))


def test_flat_tokenizer_bit_exact(text):
    tok = text.FlatPhonemeTokenizer()
    assert tok.size == GOLD["flat_size"] == 253
    assert tok.phoneme2idx == GOLD["flat_phoneme2idx"]
    assert (tok.pad_idx, tok.bos_idx, tok.eos_idx, tok.blank_idx) == (0, 1, 2, 3)
    import torch
    for s, enc, dec in zip(GOLD["sentences"], GOLD["flat_encode"], GOLD["flat_decode"]):
        assert tok(s, max_length=64) == enc, s
        assert tok.decode(torch.tensor(enc)) == dec
    assert tok("xin chào việt nam", max_length=20)[:15] == [1, 16, 86, 3, 13, 38, 248, 3, 17, 83, 252, 3, 21, 34, 2]
    for s in GOLD["flat_key_errors"]:
        with pytest.raises(KeyError):
            tok(s, max_length=64)
    batch = tok(["xin chào", "việt nam"], max_length=12)
    assert batch.shape == (2, 12) and batch[0, 0] == 1
    assert tok.create_mask(batch).sum() > 0


def test_foreign_decomposition_and_preprocess(text):
    for w, ref in zip(GOLD["foreign_words"], GOLD["foreign_decompose"]):
        assert [list(t) for t in text.decompose_foreign(w)] == ref, w
    for s, ref in zip(GOLD["preprocess_in"], GOLD["preprocess_out"]):
        assert text.preprocess_sentence(s) == ref


def test_vocab_builder_bit_exact(text):
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ann.json")
        with open(path, "w", encoding="utf-8") as f:
            json.dump(text_cases.annotations(), f, ensure_ascii=False)
        vocab = text.VocabBuilder([path]).vocab
    assert vocab == GOLD["vocab"]
    assert [list(vocab[k].items()) for k in ("onset", "rhyme", "tone")] == \
           [list(GOLD["vocab"][k].items()) for k in ("onset", "rhyme", "tone")]        # same ids, same order


def test_three_vocab_tokenizer_contract(text):
    """the reconstructed tokenizer the executor expects: shapes, specials, round trip"""
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ann.json")
        with open(path, "w", encoding="utf-8") as f:
            json.dump(text_cases.annotations(), f, ensure_ascii=False)
        tok = text.PhonemeTokenizer(vocab_path=os.path.join(td, "vocab.json"), annotation_paths=[path], max_length=32)
        tok2 = text.PhonemeTokenizer(vocab_path=os.path.join(td, "vocab.json"), annotation_paths=None, max_length=32)
    assert tok.vocab == tok2.vocab
    assert (tok.pad_id, tok.bos_id, tok.eos_id) == (2, 3, 4)
    ids = tok("phở bò")
    assert len(ids) == 32 and all(len(t) == 3 for t in ids)
    assert ids[0] == [3, 0, 0] and [4, 0, 0] in ids and ids[-1] == [2, 2, 2]
    v = tok.vocab
    assert ids[1] == [v["onset"]["ph"], v["rhyme"]["ơ"], v["tone"]["<?>"]]
    assert ids[2] == [v["onset"]["<_>"], 0, 0]
    mask = tok.create_mask(ids)
    assert len(mask) == 1 and len(mask[0]) == 32 and mask[0][0] == 0 and mask[0][-1] == 1
    assert tok.batch_decode([ids, tok("màu xanh coca")]) == ["phở bò", "màu xanh coca"]
