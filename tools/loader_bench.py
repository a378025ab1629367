"""Host-side throughput of the input pipeline (samples/s, CPU only): the reference's `PhonemeLaTrDataset` behind
`DataLoader(shuffle=True)` (when /root/reference is present: build container only) against the packed dataset +
`PinnedBatchLoader` of `phoneme-vqa_b200/data.py`, on a synthetic on-disk dataset at the real per-sample sizes
(L_ocr = 100, L_q = 30, T = 20, 3x224x224 fp32 features).   usage: python tools/loader_bench.py [n_images] [n_qa]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoneme_vqa_b200 import data, text  # noqa: E402

WORDS = ["cửa", "hàng", "bánh", "mì", "số", "12", "phở", "Hà", "Nội", "SALE", "50%", "đường", "Nguyễn", "Trãi", "café",
         "trà", "sữa", "MILK", "tea", "quán", "ăn", "ngon", "giá", "rẻ", "mở", "7h-22h", "wifi", "free", "ATM", "xăng"]


class _Enc(dict):
    __getattr__ = dict.__getitem__


class StubT5Tokenizer:
    """the three call forms the dataset uses (HF T5 tokenizer contract), 3-character pieces with hashed ids; the real
    sentencepiece model is not available offline and its cost is the same for both arms"""
    eos_token_id, pad_token_id = 1, 0

    def _text(self, t):
        return [3 + hash(w[i:i + 3]) % 32000 for w in str(t).split() if w != "<pad>" for i in range(0, len(w), 3)]

    def __call__(self, text, padding=False, max_length=None, truncation=False, is_split_into_words=False,
                 add_special_tokens=True):
        if isinstance(text, (list, tuple)):
            if is_split_into_words:
                ids = [i for w in text for i in self._text(w)]
                return _Enc(input_ids=ids, attention_mask=[1] * len(ids))
            return _Enc(input_ids=[self._text(w) for w in text])
        ids = self._text(text)[:max_length - 1] + [1]
        pad = max_length - len(ids)
        return _Enc(input_ids=ids + [0] * pad, attention_mask=[1] * len(ids) + [0] * pad)


def build(root, n_images, n_qa, seed=0):
    import pandas as pd
    rng = np.random.RandomState(seed)
    ocr_root, feat_root = os.path.join(root, "ocr"), os.path.join(root, "features")
    os.makedirs(ocr_root); os.makedirs(feat_root)
    for i in range(n_images):
        n = int(rng.randint(5, 60))
        xy = rng.uniform(0, 0.9, size=(n, 2)); wh = rng.uniform(0.001, 0.1, size=(n, 2))
        np.save(os.path.join(ocr_root, f"{i}.npy"), {"texts": [WORDS[int(k)] for k in rng.randint(0, len(WORDS), n)],
                                                      "boxes": np.concatenate([xy, xy + wh], 1)}, allow_pickle=True)
        np.save(os.path.join(feat_root, f"{i}.npy"), {"image": rng.standard_normal((1, 3, 224, 224)).astype(np.float32)},
                allow_pickle=True)
    qa = pd.DataFrame({"image_id": rng.randint(0, n_images, n_qa), "question": ["cửa hàng này bán gì ?"] * n_qa,
                       "answer": ["bánh mì hà nội"] * n_qa, "filename": ["x.jpg"] * n_qa})
    ann = os.path.join(root, "ann.json")
    with open(ann, "w", encoding="utf-8") as f:
        json.dump({"annotations": [{"question": "cửa hàng này bán gì ?", "answers": ["bánh mì hà nội"]}]}, f, ensure_ascii=False)
    return ocr_root, feat_root, qa, text.PhonemeTokenizer(None, [ann], max_length=20)


def rate(it, n_samples):
    t0 = time.perf_counter()
    seen = 0
    for b in it:
        seen += b["input_ids"].shape[0]
    assert seen == n_samples, (seen, n_samples)
    return n_samples / (time.perf_counter() - t0)


def main():
    n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_qa = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    out = {"n_images": n_images, "n_qa": n_qa, "batch": 64, "cores": os.cpu_count()}
    with tempfile.TemporaryDirectory() as tmp:
        ocr_root, feat_root, qa, ptok = build(tmp, n_images, n_qa)
        ocr_df = data.textlayout_ocr_adapt(ocr_root)
        t0 = time.perf_counter()
        ds = data.PhonemeLaTrDataset(qa, ocr_df, StubT5Tokenizer(), ptok, feat_root)
        out["packed_encode_s"] = time.perf_counter() - t0
        out["packed_per_file_features"] = rate(data.PinnedBatchLoader(ds, 64, shuffle=True), len(ds))
        ds.pack_features(os.path.join(tmp, "pack.npy"))
        rate(data.PinnedBatchLoader(ds, 64, shuffle=True), len(ds))           # page the pack in once
        out["packed_feature_pack"] = rate(data.PinnedBatchLoader(ds, 64, shuffle=True), len(ds))
        # what the device path pays on the host: gather into the rotating (pinned) buffers, no further copy
        out["packed_feature_pack_borrowed"] = rate(data.PinnedBatchLoader(ds, 64, shuffle=True, borrow=True), len(ds))
        if os.path.isdir("/root/reference/core"):
            sys.path.insert(0, "/root/reference")
            from core.data import PhonemeLaTrDataset as RefDataset
            from torch.utils.data import DataLoader
            t0 = time.perf_counter()
            ref = RefDataset(qa, ocr_df, StubT5Tokenizer(), ptok, feat_root)
            out["reference_encode_s"] = time.perf_counter() - t0
            for workers in (0, 4):
                out[f"reference_dataloader_workers{workers}"] = rate(DataLoader(ref, batch_size=64, shuffle=True,
                                                                                num_workers=workers), len(ref))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
