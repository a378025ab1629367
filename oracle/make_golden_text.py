"""Text-pipeline goldens from the REAL reference functions (build container only; needs /root/reference).

Writes tests/golden/text_golden.json: explicit outputs for ~700 words / 40 sentences and SHA-256 digests of the
outputs over ~37k generated syllables, for both is_Vietnamese variants, compose_word, the flat
PhonemeTokenizer and VocabBuilder.  The word list itself is regenerated deterministically by
`oracle.text_cases` so the fixture stays small."""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PVQA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import text_cases  # noqa: E402


def digest(items):
    h = hashlib.sha256()
    for it in items:
        h.update(json.dumps(it, ensure_ascii=False, sort_keys=True).encode("utf-8"))
        h.update(b"\n")
    return h.hexdigest()


def main():
    from core.tokenizer.modules.word_processing import is_Vietnamese as core_is_vn
    from core.tokenizer.modules.vocab_builder import VocabBuilder
    from core.tokenizer.phoneme_tokenizer import PhonemeTokenizer
    import decode.word_processing as dwp

    words = text_cases.all_words()
    core_out = [list(map(_plain, core_is_vn(w))) for w in words]
    dec_out = [list(map(_plain, dwp.is_Vietnamese(w))) for w in words]
    composed = [dwp.compose_word(*r[1]) if r[0] else None for r in dec_out]
    tok = PhonemeTokenizer()
    sentences, enc, key_errors = [], [], []
    for s in text_cases.sentences():
        try:
            e = tok(s, max_length=64)
        except KeyError:
            key_errors.append(s)          # symbol outside the 253-entry inventory: the reference raises
            continue
        sentences.append(s)
        enc.append(e)
    import torch
    dec = [tok.decode(torch.tensor(e)) for e in enc]
    foreign = [[list(t) for t in dwp.decompose_non_vietnamese_word(w)] for w in text_cases.CURATED[-40:]]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ann.json")
        with open(path, "w", encoding="utf-8") as f:
            json.dump(text_cases.annotations(), f, ensure_ascii=False)
        vocab = VocabBuilder([path]).vocab
    n_explicit = len(text_cases.CURATED) + 400
    out = {
        "n_words": len(words),
        "core_sha256": digest(core_out), "decode_sha256": digest(dec_out), "compose_sha256": digest(composed),
        "explicit_words": words[:n_explicit],
        "explicit_core": core_out[:n_explicit], "explicit_decode": dec_out[:n_explicit],
        "explicit_compose": composed[:n_explicit],
        "flat_phoneme2idx": tok.phoneme2idx, "flat_size": tok.size,
        "sentences": sentences, "flat_encode": enc, "flat_decode": dec, "flat_key_errors": key_errors,
        "foreign_words": text_cases.CURATED[-40:], "foreign_decompose": foreign,
        "preprocess_in": text_cases.RAW_SENTENCES,
        "preprocess_out": [dwp.preprocess_sentence(s) for s in text_cases.RAW_SENTENCES],
        "vocab": vocab,
    }
    with open(os.path.join(ROOT, "tests", "golden", "text_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    print("text golden:", len(words), "words; vietnamese(core) =", sum(1 for r in core_out if r[0]),
          "vocab sizes", {k: len(v) for k, v in vocab.items()})


def _plain(x):
    if isinstance(x, tuple):
        return list(x)
    return x


if __name__ == "__main__":
    main()
