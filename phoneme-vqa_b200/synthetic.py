"""Model configs and seeded synthetic batches of the BASELINE.json workloads.

No network / datasets here: weights are random-init from config and batches are synthetic,
with the tensor contract of core/data/PhonemeLaTrDataset.py:51-58 (names, dtypes, padding:
eos box [1000]*6, pad box [0]*6, pad token 0, float masks)."""
from __future__ import annotations

import torch

PHONEME_VOCAB = (84, 187, 7)     # full-inventory (onset, rhyme, tone) sizes, SURVEY §8
PAD_ID, BOS_ID, EOS_ID = 2, 3, 4  # onset-vocabulary specials of vocab_builder.py:15-31


def t5_config(size="base", vocab_size=36096, num_decoder_layers=4, n_head=None, vit_config=None, **kw):
    """T5 dims of VietAI/vit5-{small,base,large} + the 4 extra keys CustomizedLaTr_config.build adds
    (core/model/PhonemeLaTr.py:6-15).  n_head follows SURVEY §8d (YAML's 12 only divides d=768)."""
    from transformers import T5Config
    dims = {"small": dict(d_model=512, d_kv=64, num_heads=8, d_ff=2048, num_layers=6),
            "base": dict(d_model=768, d_kv=64, num_heads=12, d_ff=3072, num_layers=12),
            "large": dict(d_model=1024, d_kv=64, num_heads=16, d_ff=4096, num_layers=24)}[size]
    cfg = T5Config(vocab_size=vocab_size, dropout_rate=0.1, feed_forward_proj="relu", decoder_start_token_id=0, **dims)
    cfg.update({"max_2d_position_embeddings": 1024, "vit_model": "google/vit-base-patch16-224-in21k",
                "num_decoder_layers": num_decoder_layers, "n_head": n_head or dims["num_heads"],
                "random_init": True, "vit_config": vit_config})
    cfg.update(kw)
    return cfg


def phoneme_latr_batch(B, vocab_size, T=127, L_ocr=100, L_q=30, V_sub=PHONEME_VOCAB, seed=1234, image=224,
                       device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(min(20, L_ocr - 1), min(100, L_ocr), (B,), generator=g)
    pos = torch.arange(L_ocr)[None, :]
    x0 = torch.randint(0, 901, (B, L_ocr), generator=g)
    y0 = torch.randint(0, 901, (B, L_ocr), generator=g)
    w = torch.randint(1, 101, (B, L_ocr), generator=g)
    h = torch.randint(1, 101, (B, L_ocr), generator=g)
    coords = torch.stack([x0, y0, x0 + w, y0 + h, w, h], dim=-1)
    valid = (pos < n[:, None])
    iseos = (pos == n[:, None])
    coords = torch.where(valid[..., None], coords, torch.zeros_like(coords))
    coords = torch.where(iseos[..., None], torch.full_like(coords, 1000), coords)
    ocr = torch.randint(3, vocab_size, (B, L_ocr), generator=g)
    ocr = torch.where(valid, ocr, torch.zeros_like(ocr))
    ocr = torch.where(iseos, torch.ones_like(ocr), ocr)
    om = (pos <= n[:, None]).float()
    nq = torch.randint(min(8, L_q), L_q + 1, (B,), generator=g)
    qpos = torch.arange(L_q)[None, :]
    q = torch.randint(3, vocab_size, (B, L_q), generator=g)
    q = torch.where(qpos < nq[:, None] - 1, q, torch.zeros_like(q))
    q = torch.where(qpos == nq[:, None] - 1, torch.ones_like(q), q)
    qm = (qpos < nq[:, None]).float()
    ln = torch.randint(min(4, T), min(40, T) + 1, (B,), generator=g)
    tpos = torch.arange(T + 1)[None, :]
    lab = torch.stack([torch.randint(5, V_sub[0], (B, T + 1), generator=g),
                       torch.randint(2, V_sub[1], (B, T + 1), generator=g),
                       torch.randint(0, V_sub[2], (B, T + 1), generator=g)], dim=-1)
    bos = torch.tensor([BOS_ID, 0, 0])
    eos = torch.tensor([EOS_ID, 0, 0])
    lab = torch.where((tpos == 0)[..., None], bos, lab)
    lab = torch.where((tpos == ln[:, None])[..., None], eos, lab)
    lab = torch.where((tpos > ln[:, None])[..., None], torch.full_like(lab, PAD_ID), lab)
    lmask = (tpos > ln[:, None]).float()          # 1.0 = pad (float(create_mask), PhonemeLaTrDataset.py:55)
    pix = torch.randn(B, 3, image, image, generator=g)
    batch = {"pixel_values": pix, "coordinates": coords, "input_ids": q, "src_attention_mask": qm,
             "label_ids": lab, "label_attention_mask": lmask, "tokenized_ocr": ocr, "ocr_attention_mask": om}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def batch_bytes(batch) -> int:
    return sum(v.numel() * v.element_size() for v in batch.values())


def latr_batch(B, vocab_size, T=127, L_ocr=100, L_q=30, seed=1234, image=224, device="cpu", pin=False):
    """plain LaTr labels: "<pad> " + answer sub-tokens + eos (id 1), pad id 0; attention mask 1 = valid (int64)
    (core/data/LaTrDataset.py answer encoding)."""
    batch = phoneme_latr_batch(B, vocab_size, T=T, L_ocr=L_ocr, L_q=L_q, seed=seed, image=image)
    g = torch.Generator().manual_seed(seed + 17)
    ln = torch.randint(min(4, T), min(40, T) + 1, (B,), generator=g)
    tpos = torch.arange(T + 1)[None, :]
    lab = torch.randint(3, vocab_size, (B, T + 1), generator=g)
    lab = torch.where(tpos == 0, torch.zeros_like(lab), lab)
    lab = torch.where(tpos == ln[:, None], torch.ones_like(lab), lab)
    lab = torch.where(tpos > ln[:, None], torch.zeros_like(lab), lab)
    batch["label_ids"] = lab
    batch["label_attention_mask"] = (tpos <= ln[:, None]).long()
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def phoneme_prestu_batch(B, vocab_size, T=127, L_in=130, V_sub=PHONEME_VOCAB, seed=1234, image=224, device="cpu",
                         pin=False):
    """PhonemePreSTUDataset fields (core/data/PhonemePreSTUDataset.py): question and OCR text share one
    `input_ids` sequence of max_q_length + max_ocr_length = 130 (config/phonemeprestu.yaml:37-38), no boxes."""
    b = phoneme_latr_batch(B, vocab_size, T=T, L_ocr=4, L_q=L_in, V_sub=V_sub, seed=seed, image=image)
    batch = {k: b[k] for k in ("pixel_values", "input_ids", "src_attention_mask", "label_ids", "label_attention_mask")}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def phoneme_prestu_loss(model, b, ignore_index=PAD_ID):
    """step body of core/executor/PhonemePreSTU_Executor.py:150-185 on the fused head"""
    labels = b["label_ids"]
    return model.forward_loss(b["pixel_values"], b["input_ids"], labels[:, :-1], b["src_attention_mask"],
                              b["label_attention_mask"][:, :-1], targets=labels[:, 1:], ignore_index=ignore_index)


SAL_FIELDS = ("input_ids", "src_attention_mask", "label_ids", "shifted_right_label_ids", "label_attention_mask",
              "tokenized_ocr", "ocr_attention_mask", "ocr_coordinates", "ocr_features", "tokenized_obj",
              "obj_attention_mask", "obj_coordinates", "obj_features")


def phoneme_sal_batch(B, vocab_size, T=39, L_q=80, L_ocr=256, L_obj=128, ocr_hidden=512, obj_hidden=2048,
                      tgt_vocab=253, seed=1234, device="cpu", pin=False):
    """PhonemeSaLDataset fields (core/data/PhonemeSaLDataset.py:78-92): float masks, boxes in
    [0, 0.99), region features, flat 253-phoneme labels (pad 0, bos 1, eos 2), bool label mask = pad positions."""
    g = torch.Generator().manual_seed(seed)

    def toks(L, lo):
        n = torch.randint(lo, L, (B,), generator=g)
        pos = torch.arange(L)[None, :]
        ids = torch.randint(3, vocab_size, (B, L), generator=g)
        ids = torch.where(pos < n[:, None], ids, torch.zeros_like(ids))
        ids = torch.where(pos == n[:, None], torch.ones_like(ids), ids)
        return ids, (pos <= n[:, None]).float()

    def boxes(L):
        xy = torch.rand(B, L, 2, generator=g) * 0.8
        return torch.cat([xy, xy + torch.rand(B, L, 2, generator=g) * 0.19], dim=-1)

    q, qm = toks(L_q, min(8, L_q - 1))
    ocr, om = toks(L_ocr, min(20, L_ocr - 1))
    obj, bm = toks(L_obj, min(10, L_obj - 1))
    ln = torch.randint(min(4, T), T, (B,), generator=g)
    tpos = torch.arange(T + 1)[None, :]
    lab = torch.randint(4, tgt_vocab, (B, T + 1), generator=g)
    lab = torch.where(tpos == 0, torch.ones_like(lab), lab)
    lab = torch.where(tpos == ln[:, None], torch.full_like(lab, 2), lab)
    lab = torch.where(tpos > ln[:, None], torch.zeros_like(lab), lab)
    batch = {"input_ids": q, "src_attention_mask": qm, "label_ids": lab[:, :-1].contiguous(),
             "shifted_right_label_ids": lab[:, 1:].contiguous(), "label_attention_mask": lab[:, :-1] == 0,
             "tokenized_ocr": ocr, "ocr_attention_mask": om, "ocr_coordinates": boxes(L_ocr),
             "ocr_features": torch.randn(B, L_ocr, ocr_hidden, generator=g),
             "tokenized_obj": obj, "obj_attention_mask": bm, "obj_coordinates": boxes(L_obj),
             "obj_features": torch.randn(B, L_obj, obj_hidden, generator=g)}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def phoneme_sal_loss(L_ocr, L_q):
    """step body of core/executor/PhonemeSaL_Executor.py (the model returns (logits, loss) itself)"""
    def loss_fn(model, b):
        return model(*[b[k] for k in SAL_FIELDS], L_ocr, L_q)[1]
    return loss_fn
