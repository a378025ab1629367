"""Vietnamese phoneme text pipeline: words -> (onset, rhyme, tone) -> index tensors.

Bit-exact restatement (table-driven, no code shared) of the reference's rule-based pipeline:

  * ``analyse(word, "core")``    == core/tokenizer/modules/word_processing.py:121-288  is_Vietnamese (3-tuple)
  * ``analyse(word, "decode")``  == decode/word_processing.py:97-247                    is_Vietnamese (5-tuple)
  * ``compose_word``             == decode/word_processing.py:274-317
  * ``decompose_foreign``        == decode/word_processing.py:249-272  (incl. its ``"ê" "i"`` literal quirk)
  * ``preprocess_sentence``      == decode/word_processing.py:319-334
  * ``FlatPhonemeTokenizer``     == core/tokenizer/phoneme_tokenizer.py:5-177 (253-symbol flat vocabulary)
  * ``VocabBuilder``             == core/tokenizer/modules/vocab_builder.py:11-113
  * ``PhonemeTokenizer``         the 3-vocabulary tokenizer the executor and dataset call
                                 (core/executor/PhonemeLaTr_Executor.py:282-287, core/data/PhonemeLaTrDataset.py:47,85)
                                 but the snapshot does not define — a documented reconstruction.

Everything here is CPU string/integer work done once at dataset build time; it is pinned against the
reference on ~37k generated syllables + hand-picked words by tests/test_text_golden.py.
"""
from __future__ import annotations

import json
import re
import string
import unicodedata

# ---------------------------------------------------------------------------------
# inventories
# ---------------------------------------------------------------------------------
_TONE_MARKS = ("̀", "́", "̃", "̉", "̣")          # huyền sắc ngã hỏi nặng
_TONE_NAMES = {"core": ("<`>", "</>", "<~>", "<?>", "<.>"),
               "decode": ("<huyền>", "<sắc>", "<ngã>", "<hỏi>", "<nặng>")}
# longest-match-first order matters ("ngh" before "ng" before "n", "gi"/"gh" before "g")
ONSETS = ("ngh", "tr", "th", "ph", "nh", "ng", "kh", "gi", "gh", "ch", "q", "đ", "x", "v", "t", "s", "r", "n", "m",
          "l", "k", "h", "g", "d", "c", "b")
NUCLEI = ("oo", "ươ", "ưa", "uô", "ua", "iê", "yê", "ia", "ya", "e", "ê", "u", "ư", "ô", "i", "y", "o", "ơ", "â", "a",
          "o", "ă")
CODAS = frozenset(("ng", "nh", "ch", "u", "n", "o", "p", "c", "m", "y", "i", "t"))
_U_MEDIAL_BEFORE = ("ê", "y", "ơ", "a", "â", "ya")
_O_MEDIAL_BEFORE = ("oa", "oă", "oe")
_SINGLE_VOWELS = frozenset(n for n in NUCLEI if len(n) == 1)
_LEADING = re.compile(r"[a-zA-Zăâđưôơê]")
_SPECIAL = {"gin": "giin", "giêng": "giiêng", "giêt": "giiêt", "giêc": "giiêc", "gi": "gii"}
_FOREIGN_ONSETS = frozenset(("m", "b", "v", "t", "đ", "n", "x", "s", "l", "h", "r", "g", "d", "k", "q", "c", "ph", "th",
                             "nh", "tr", "ch", "kh", "gh", "gi", "ng", "ngh"))

_FRONT = frozenset(("i", "y", "e", "ê", "iê", "yê", "ia", "ya"))
_FRONT_NO_Y = frozenset(("i", "e", "ê", "iê"))
_FRONT_NGH = frozenset(("i", "e", "ê", "iê", "yê", "ia", "ya"))


def strip_tone(word: str):
    """NFD-decompose, remove the (last) tone mark, NFC-recompose.  Returns (tone index | None, base)."""
    tone, rest = None, []
    for ch in unicodedata.normalize("NFD", word):
        if ch in _TONE_MARKS:
            tone = _TONE_MARKS.index(ch)
        else:
            rest.append(ch)
    return tone, unicodedata.normalize("NFC", "".join(rest))


def split_syllable(word: str):
    """(onset, medial, nucleus, coda) by greedy left-to-right matching; None for an absent slot."""
    onset = next((o for o in ONSETS if word.startswith(o)), None)
    if onset is not None and onset != "q":
        word = word[len(onset):]
    medial = None
    if word.startswith("q"):
        medial, word = "u", word.removeprefix("qu")
    elif word.startswith(_O_MEDIAL_BEFORE):
        medial, word = "o", word[1:]
    elif not (word.startswith("ua") or word.startswith("uô")):
        if any(word.startswith("u" + n) for n in _U_MEDIAL_BEFORE):
            medial, word = "u", word[1:]
    nucleus = next((n for n in NUCLEI if word.startswith(n)), None)
    if nucleus is not None:
        word = word[len(nucleus):]
    coda = word if word in CODAS else None
    return onset, medial, nucleus, coda


# rule tables: a syllable is rejected when ANY predicate holds.  (o, m, n, c) = onset, medial, nucleus, coda
_RULES_COMMON = (
    lambda o, m, n, c: o == "k" and m is None and n not in _FRONT,
    lambda o, m, n, c: o == "c" and m is None and n in _FRONT,
    lambda o, m, n, c: o == "q" and m != "u",
    lambda o, m, n, c: o == "gh" and m is None and n not in _FRONT_NO_Y,
    lambda o, m, n, c: o == "g" and m is None and n in _FRONT_NO_Y,
    lambda o, m, n, c: o == "ngh" and m is None and n not in _FRONT_NGH,
    lambda o, m, n, c: o == "ng" and m is None and n in _FRONT_NGH,
    lambda o, m, n, c: m == "o" and n not in ("a", "ă", "e"),
    lambda o, m, n, c: m == "u" and n not in ("yê", "ya", "e", "ê", "y", "ơ", "ô", "a", "â", "ă"),
    lambda o, m, n, c: n == "oo" and c not in ("ng", "c"),
    lambda o, m, n, c: n in ("ua", "ia", "ya") and c is not None,
    lambda o, m, n, c: n in ("ua", "uô") and c == "ph",
    lambda o, m, n, c: n in ("yê", "iê", "ă", "â") and c is None,
    lambda o, m, n, c: m == "o" and n in ("iê", "yê", "ia", "ya"),
    lambda o, m, n, c: m is not None and n in ("u", "oo", "o", "ua", "uô", "ươ", "ưa", "ư"),
    lambda o, m, n, c: m is not None and n in ("i", "e", "ê", "ia", "ya", "iê", "yê") and c in ("m", "ph"),
    lambda o, m, n, c: c == "o" and n not in ("a", "e"),
    lambda o, m, n, c: c == "y" and n not in ("a", "â"),
    lambda o, m, n, c: c == "i" and n in ("ă", "â", "i", "e", "iê", "yê", "ia", "ya"),
    lambda o, m, n, c: c == "nh" and n not in ("a", "i", "y", "ê"),
    lambda o, m, n, c: c == "ng" and n not in ("a", "o", "ô", "u", "ư", "e", "iê", "ươ", "â", "ă", "uô", "oo"),
    lambda o, m, n, c: c == "ch" and n not in ("i", "a", "ê", "y"),
    lambda o, m, n, c: c == "c" and n in ("i", "ê", "e", "ơ"),
    lambda o, m, n, c: n == c,
)
# extra checks of the decode/ variant; the first three run BEFORE the re-assembly test there, which does not
# change the outcome because every path rejects.
_RULES_DECODE_ONLY = (
    lambda o, m, n, c: n in ("oo", "ươ", "uô", "iê", "yê") and c is None,
    lambda o, m, n, c: n == "ya" and m is None,
    lambda o, m, n, c: n == "y" and c is not None,
    lambda o, m, n, c: o in ("r", "gi") and m is not None,
    lambda o, m, n, c: c == "u" and n in ("i", "e", "ơ", "o", "ô", "y", "ia", "ya", "oo", "ưa", "ă"),
)


def _foreign_triple(word: str):
    d = unicodedata.normalize("NFD", word)
    return (d, "", "") if d in _FOREIGN_ONSETS else ("", "", d)


def analyse(word: str, variant: str = "core"):
    """The reference's two ``is_Vietnamese`` functions.

    variant "core":   (True, (onset|None, rhyme, tone|None))  or (False, (onset, "", rest))
    variant "decode": (True, (onset, medial, nucleus, coda, tone)) or (False, None)
    """
    tone_idx, base = strip_tone(word)
    tone = None if tone_idx is None else _TONE_NAMES[variant][tone_idx]
    core = variant == "core"

    def reject(w):
        return (False, _foreign_triple(w)) if core else (False, None)

    if not _LEADING.match(base):
        return reject(base)
    base = _SPECIAL.get(base, base)
    # at most two vowel clusters may START after the first character
    starts, prev = 0, base[0] in _SINGLE_VOWELS
    for ch in base[1:]:
        cur = ch in _SINGLE_VOWELS
        if cur and not prev:
            starts += 1
            if starts > 2:
                return reject(base)
        prev = cur
    o, m, n, c = split_syllable(base)
    if n is None:
        return reject(base)
    if "".join(x for x in (o, m, n, c) if x is not None) != base:
        return reject(base)
    rules = _RULES_COMMON if core else _RULES_COMMON + _RULES_DECODE_ONLY
    if any(rule(o, m, n, c) for rule in rules):
        return reject(base)
    if core:
        return True, (o, "".join(x for x in (m, n, c) if x), tone)
    return True, (o, m, n, c, tone)


def decompose_foreign(word: str):
    """decode/word_processing.py:249-272 — one 5-tuple per character.  The reference's vowel list has a
    missing comma ("ê" "i" == "êi"), so neither "ê" nor "i" counts as a vowel there; reproduced."""
    vowels = ("a", "ă", "â", "e", "êi", "o", "ô", "ơ", "u", "ư")
    out = []
    for ch in word:
        tone_idx, base = strip_tone(ch)
        tone = None if tone_idx is None else _TONE_NAMES["decode"][tone_idx]
        out.append((None, None, base, None, tone) if base in vowels else (base, None, None, None, tone))
    return out


def compose_word(onset, medial, nucleus, coda, tone):
    """decode/word_processing.py:274-317 — inverse of the decode-variant analysis."""
    if nucleus is None:
        return onset
    names = _TONE_NAMES["decode"]
    if tone in names:
        mark = _TONE_MARKS[names.index(tone)]
        if onset != "q" and medial is not None and coda is None and nucleus not in ("ơ", "ê"):
            medial += mark
        elif coda is None:
            nucleus = nucleus[0] + mark + nucleus[1:]
        else:
            nucleus = nucleus + mark
    elif not (tone == "<blank>" or tone is None):
        raise AssertionError(f"Received tone {tone}")
    word = "".join(x for x in (onset, medial, nucleus, coda) if x)
    if "gii" in word:
        word = word.replace("gii", "gi")
    return unicodedata.normalize("NFC", word)


def preprocess_sentence(sentence: str) -> str:
    s = sentence.lower()
    for a, b in (("&", " và "), ("_", ""), ("#", ""), ("|", ""), ("~", ""), (";", " , "), ("/", " / "),
                 ("\\", " / "), ("=", " bằng ")):
        s = s.replace(a, b)
    return " ".join(s.split())


# ---------------------------------------------------------------------------------
# flat 253-symbol tokenizer (core/tokenizer/phoneme_tokenizer.py)
# ---------------------------------------------------------------------------------
_FLAT_RHYMES = (
    "a ac ach ai am an ang anh ao ap at ay au "
    "ă ăc ăm ăn ăng ăp ăt "
    "â âc âm ân âng âp ât âu ây "
    "e ec em en eng eo ep et "
    "ê êch êm ên ênh êp êt êu "
    "i ia ich iêc iêm iên iêng iêp iêt iêu im in inh ip it iu "
    "o oa oac oach oai oam oan oang oanh oao oap oat oay oăc oăm oăn oăng oăt oc oe oen oeo oet oi om on ong ooc "
    "oong op ot "
    "ô ôc ôi ôm ôn ông ôp ôt "
    "ơ ơi ơm ơn ơp ơt "
    "u ua uân uâng uât uây uc uê uêch uênh ui um un ung uơ uôc uôi uôm uôn uông uôt up ut uy uya uych uyên uyêt "
    "uyn uynh uyp uyt uyu uach uai uan uang uanh uao uat uau uay uăc uăm uăn uăng uăp uăt uâc uoang ue uen ueo uet "
    "uên uêt uêu uơi "
    "ư ưa ưc ưi ưng ươc ươi ươm ươn ương ươp ươt ươu ưt ưu "
    "y yêm yên yêng yêt yêu "
    "? , . - / ! @ ( ) : % \" * ' + $ < > "
    "0 1 2 3 4 5 6 7 8 9 "
    "w f z j p").split()


class FlatPhonemeTokenizer:
    """253 ids: 4 specials, 26 onsets, 218 rhymes/symbols, 5 tones; `<blank>` separates words."""

    def __init__(self):
        self.pad_token, self.bos_token, self.eos_token, self.blank_token = "<pad>", "<bos>", "<eos>", "<blank>"
        self.special_tokens = [self.pad_token, self.bos_token, self.eos_token, self.blank_token]
        symbols = self.special_tokens + list(ONSETS) + _FLAT_RHYMES + list(_TONE_NAMES["decode"])
        self.phoneme2idx = {s: i for i, s in enumerate(symbols)}      # later duplicates win, like a dict literal
        self.idx2phoneme = {i: s for s, i in self.phoneme2idx.items()}
        self.pad_idx, self.bos_idx, self.eos_idx, self.blank_idx = (self.phoneme2idx[t] for t in self.special_tokens)

    @property
    def size(self) -> int:
        return len(self.phoneme2idx)

    def encode(self, sentence: str, max_length: int):
        comps = []
        for word in sentence.split():
            ok, parts = analyse(word, "decode")
            if ok:
                comps.append(parts)
            else:
                comps.extend(decompose_foreign(word))
        ids = []
        for onset, medial, nucleus, coda, tone in comps:
            rhyme = compose_word(None, medial, nucleus, coda, None)
            for sym in (onset, rhyme, tone):
                if sym:
                    ids.append(self.phoneme2idx[sym])
            ids.append(self.blank_idx)
        ids = [self.bos_idx] + ids[:-1] + [self.eos_idx]
        if len(ids) < max_length:
            ids.extend([self.pad_idx] * (max_length - len(ids)))
        else:
            ids = ids[:max_length]
        return ids

    def batch_encode(self, sentences, max_length):
        import torch
        return torch.tensor([self.encode(s.lower(), max_length) for s in sentences])

    def decode(self, ids) -> str:
        ids = ids.long().tolist() if hasattr(ids, "long") else list(ids)
        out = []
        for i in ids:
            sym = self.idx2phoneme[i]
            out.append(" " if sym == self.blank_token else sym)
        text = "".join(s for s in out if s not in self.special_tokens)
        return " ".join(text.split())

    def batch_decode(self, matrix):
        return [self.decode(row) for row in matrix]

    def __call__(self, sentences, max_length=30):
        if isinstance(sentences, str):
            return self.encode(sentences.lower(), max_length=max_length)
        if isinstance(sentences, list):
            return self.batch_encode(sentences, max_length=max_length)

    def create_mask(self, ids):
        return (ids == self.pad_idx).bool()


# ---------------------------------------------------------------------------------
# three-vocabulary builder (core/tokenizer/modules/vocab_builder.py)
# ---------------------------------------------------------------------------------
class VocabBuilder:
    def __init__(self, annotation_paths=None):
        self.annotation_paths = annotation_paths
        self.vocab = {"onset": {"none": 0, "<_>": 1, "<pad>": 2, "<bos>": 3, "<eos>": 4},
                      "rhyme": {"none": 0, "<pad>": 1},
                      "tone": {"none": 0, "<pad>": 1}}
        self.word_counts = self.create_vocab()

    def _add(self, kind, key):
        table = self.vocab[kind]
        if key not in table:
            table[key] = len(table)

    def create_vocab(self):
        printable = string.ascii_lowercase + string.digits + string.punctuation
        for path in self.annotation_paths:
            with open(path, "r", encoding="utf-8") as f:
                annotations = json.load(f).get("annotations", [])
            for ann in annotations:
                for field in ("question", "answers"):
                    if field not in ann:
                        continue
                    text = ann[field] if isinstance(ann[field], str) else ann[field][0]
                    for word in text.split():
                        word = word.lower()
                        ok, (onset, rhyme, tone) = analyse(word, "core")
                        if ok:
                            self._add("onset", onset.lower() if onset else "none")
                            self._add("rhyme", rhyme.lower() if rhyme else "none")
                            self._add("tone", tone.lower() if tone else "none")
                        else:
                            for ch in word:
                                if ch.islower():
                                    self._add("onset", ch)
                            for ch in printable:
                                self._add("onset", ch)
        return self.vocab

    def save_vocab(self, output_path: str):
        with open(output_path, "w", encoding="utf-8") as f:
            json.dump(self.vocab, f, ensure_ascii=False, indent=4)


# ---------------------------------------------------------------------------------
# the tokenizer the PhonemeLaTr executor/dataset expect (RECONSTRUCTION — see module docstring)
# ---------------------------------------------------------------------------------
class PhonemeTokenizer:
    """``PhonemeTokenizer(vocab_path=..., annotation_paths=...)`` with ``.vocab['onset'|'rhyme'|'tone']``,
    ``.pad_id/.bos_id/.eos_id``, ``__call__(text) -> [[onset, rhyme, tone], ...]``, ``create_mask`` and
    ``batch_decode`` — the interface used at core/executor/PhonemeLaTr_Executor.py:45-57,282-287 and
    core/data/PhonemeLaTrDataset.py:47,85.

    Encoding rules (written down because the snapshot has no implementation):
      * a Vietnamese word  -> one triple (onset|'none', rhyme|'none', tone|'none') looked up in the three vocabularies;
      * any other word     -> one triple per character: the character in the ONSET vocabulary
                              (vocab_builder.py:94-107 puts foreign characters there), rhyme = tone = 'none';
      * `<_>` triple between words; `<bos>` first, `<eos>` last, both (id, 0, 0) like model.generate's start
        symbol (core/model/PhonemeLaTr.py:191); padded with the onset `<pad>` id in all three columns
        (one `ignore_index` for the three losses, PhonemeLaTr_Executor.py:263-265).
      * characters/phonemes missing from a vocabulary map to 'none' (0).
    """

    def __init__(self, vocab_path=None, annotation_paths=None, max_length=128):
        import os
        if vocab_path is not None and os.path.isfile(vocab_path):
            with open(vocab_path, "r", encoding="utf-8") as f:
                self.vocab = json.load(f)
        else:
            builder = VocabBuilder(annotation_paths or [])
            self.vocab = builder.vocab
            if vocab_path is not None:
                builder.save_vocab(vocab_path)
        self.max_length = max_length
        on = self.vocab["onset"]
        self.pad_id, self.bos_id, self.eos_id, self.space_id = on["<pad>"], on["<bos>"], on["<eos>"], on["<_>"]
        self.inverse = {k: {i: s for s, i in v.items()} for k, v in self.vocab.items()}

    def encode_word(self, word):
        ok, (onset, rhyme, tone) = analyse(word, "core")
        v = self.vocab
        if ok:
            return [[v["onset"].get(onset or "none", 0), v["rhyme"].get(rhyme or "none", 0),
                     v["tone"].get(tone or "none", 0)]]
        return [[v["onset"].get(ch, 0), 0, 0] for ch in word]

    def __call__(self, text, max_length=None):
        max_length = max_length or self.max_length
        triples = [[self.bos_id, 0, 0]]
        words = text.lower().split()
        for wi, word in enumerate(words):
            if wi:
                triples.append([self.space_id, 0, 0])
            triples.extend(self.encode_word(word))
        triples = triples[: max_length - 1] + [[self.eos_id, 0, 0]]
        triples += [[self.pad_id] * 3] * (max_length - len(triples))
        return triples

    def create_mask(self, triples):
        """1 = pad (the dataset casts this to float; nn.MultiheadAttention then ADDS it — SURVEY D14)."""
        return [[1 if t[0] == self.pad_id else 0 for t in triples]]

    def decode(self, triples) -> str:
        words, cur = [], ""
        for onset, rhyme, tone in triples:
            if onset in (self.bos_id, self.pad_id):
                continue
            if onset == self.eos_id:
                break
            if onset == self.space_id:
                words.append(cur)
                cur = ""
                continue
            o = self.inverse["onset"].get(onset, "none")
            r = self.inverse["rhyme"].get(rhyme, "none")
            t = self.inverse["tone"].get(tone, "none")
            if r == "none":
                cur += "" if o == "none" else o
            else:
                marks = dict(zip(_TONE_NAMES["core"], _TONE_MARKS))
                syl = ("" if o == "none" else o) + r
                if t in marks:
                    syl = _place_tone(syl, len("" if o == "none" else o), marks[t])
                if syl.startswith("gii"):
                    syl = "gi" + syl[3:]
                cur += unicodedata.normalize("NFC", syl)
        words.append(cur)
        return " ".join(w for w in words if w)

    def batch_decode(self, batch):
        return [self.decode(t) for t in batch]


def _place_tone(syllable: str, onset_len: int, mark: str) -> str:
    """put the tone mark on the rhyme through the decode-variant composer"""
    ok, parts = analyse(syllable, "decode")
    if ok:
        o, m, n, c, _ = parts
        names = dict(zip(_TONE_MARKS, _TONE_NAMES["decode"]))
        return compose_word(o, m, n, c, names[mark])
    return syllable[: onset_len + 1] + mark + syllable[onset_len + 1:]
