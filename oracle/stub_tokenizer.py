"""TEST INFRASTRUCTURE ONLY (see oracle/__init__ docstring): a deterministic stand-in for the HF T5 tokenizer.

The reference datasets (core/data/PhonemeLaTrDataset.py:103-151) call the tokenizer in three ways; the real
`google-t5/t5-base` sentencepiece model is not available offline, so the golden vectors for the INPUT PIPELINE are
generated with this stub on both sides (reference dataset class and the product's `data.py`).  It reproduces the call
contract the dataset code relies on, not T5's vocabulary:

  tok(text, padding='max_length', max_length=L, truncation=True)        -> {'input_ids': [L], 'attention_mask': [L]}
  tok(words, is_split_into_words=True,  add_special_tokens=False)       -> {'input_ids': flat list over all words}
  tok(words, is_split_into_words=False, add_special_tokens=False)       -> .input_ids = one id list per word
  tok.eos_token_id == 1, tok.pad_token_id == 0

A "word" is split into pieces of at most 3 characters; a piece's id is a stable hash into [3, vocab_size).
"""
import zlib


class _Encoding(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class StubT5Tokenizer:
    eos_token_id = 1
    pad_token_id = 0

    def __init__(self, vocab_size=32100):
        self.vocab_size = vocab_size

    def _piece_id(self, piece):
        return 3 + zlib.crc32(piece.encode("utf-8")) % (self.vocab_size - 3)

    def _word(self, w):
        w = str(w)
        if w == "<pad>":
            return [self.pad_token_id]
        return [self._piece_id(w[i:i + 3]) for i in range(0, len(w), 3)]

    def _text(self, text):
        ids = []
        for w in str(text).split():
            ids += self._word(w)
        return ids

    def __call__(self, text, padding=False, max_length=None, truncation=False, is_split_into_words=False,
                 add_special_tokens=True):
        if isinstance(text, (list, tuple)):
            if is_split_into_words:                       # one sequence given as words
                ids = []
                for w in text:
                    ids += self._text(w)
                if add_special_tokens:
                    ids.append(self.eos_token_id)
                return _Encoding(input_ids=ids, attention_mask=[1] * len(ids))
            seqs = [self._text(t) + ([self.eos_token_id] if add_special_tokens else []) for t in text]
            return _Encoding(input_ids=seqs, attention_mask=[[1] * len(s) for s in seqs])
        ids = self._text(text)
        if add_special_tokens:
            if truncation and max_length is not None:
                ids = ids[:max_length - 1]
            ids.append(self.eos_token_id)
        elif truncation and max_length is not None:
            ids = ids[:max_length]
        mask = [1] * len(ids)
        if padding == "max_length" and max_length is not None:
            n = max_length - len(ids)
            ids = ids + [self.pad_token_id] * n
            mask = mask + [0] * n
        return _Encoding(input_ids=ids, attention_mask=mask)


class StubFlatTokenizer:
    """call contract of the reference's char / byte / BPE decode tokenizers (core/tokenizer/byte_tokenizer.py:1-46):
    `tok(text, max_length=None)` -> [bos] + ids + [eos] + [pad]*, attribute `pad_id`.  UTF-8 bytes as ids."""
    pad_id, bos_id, eos_id = 256, 257, 258

    def __call__(self, text, max_length=None, padding=True, add_special_tokens=True):
        ids = list(text.encode("utf-8"))
        if max_length is not None:
            ids = ids[:max_length - 2]
        ids = [self.bos_id] + ids + [self.eos_id]
        if max_length is not None and padding:
            ids += [self.pad_id] * (max_length - len(ids))
        return ids
