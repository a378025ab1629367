"""CPU restatement of the reference's arithmetic on the hot path (TEST INFRASTRUCTURE ONLY).

This file is the oracle for the parity tests: plain PyTorch-on-CPU / numpy restatements
of what hieunghia-pat/phoneme-VQA computes, op by op, each citing the reference
file:line it follows (paths relative to the reference root).  Nothing under
``phoneme-vqa_b200/`` (the product) imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``oracle/make_golden.py`` (imports /root/reference read-only) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# -- core/model/PhonemeLaTr.py:33-44  SpatialModule.forward ---------------------------
def spatial_module(coordinates: torch.Tensor, tables: list[torch.Tensor]) -> torch.Tensor:
    """tables in coordinate-column order: top_left_x, top_left_y, bottom_right_x,
    bottom_right_y, width_emb, height_emb.  Sum is left-associated as in the reference."""
    feats = [F.embedding(coordinates[:, :, t], tables[t]) for t in range(6)]
    out = feats[0] + feats[1]
    for f in feats[2:]:
        out = out + f
    return out


# -- core/model/PhonemeLaTr.py:219-231  _calculate_embedding (after ViT+projector) ----
def calculate_embedding(img_feat, coordinates, tokenized_ocr, input_ids, ocr_attention_mask,
                        src_attention_mask, shared, layout_tables):
    parts, masks = [], []
    if img_feat is not None:
        parts.append(img_feat)
        masks.append(torch.ones(img_feat.shape[:2]))
    if tokenized_ocr is not None:
        ocr_feat = F.embedding(tokenized_ocr, shared)
        parts.append(ocr_feat + spatial_module(coordinates, layout_tables))
        masks.append(ocr_attention_mask.float())
    if input_ids is not None:
        parts.append(F.embedding(input_ids, shared))
        masks.append(src_attention_mask.float())
    return torch.cat(parts, dim=1), torch.cat(masks, dim=1)


# -- PhonoLaTr/modules.py:47-63 (3-table form the call site core/model/PhonemeLaTr.py:72-78 expects)
def phoneme_embedding(labels, onset, rhyme, tone):
    return torch.cat((F.embedding(labels[:, :, 0], onset),
                      F.embedding(labels[:, :, 1], rhyme),
                      F.embedding(labels[:, :, 2], tone)), dim=-1)


# -- core/model/modules/transformer_utils.py:12-21  positional table --------------------
def sinusoidal_table(emb_size: int, maxlen: int = 5000) -> torch.Tensor:
    den = torch.exp(-torch.arange(0, emb_size, 2) * math.log(10000) / emb_size)
    pos = torch.arange(0, maxlen).reshape(maxlen, 1)
    pe = torch.zeros((maxlen, emb_size))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.unsqueeze(0)


# -- core/model/modules/transformer_utils.py:23-25 (dropout off) ------------------------
def positional_encoding(x, pe):
    return x + pe[:, : x.size(1)]


# -- transformers T5Attention._relative_position_bucket (modeling_t5.py:190-235) --------
def t5_relative_bucket(rel: np.ndarray, bidirectional=True, num_buckets=32, max_distance=128) -> np.ndarray:
    """Integer restatement; float32 log exactly as torch does it (torch.log on float32)."""
    rel_t = torch.as_tensor(rel, dtype=torch.long)
    buckets = torch.zeros_like(rel_t)
    nb = num_buckets
    if bidirectional:
        nb //= 2
        buckets = buckets + (rel_t > 0).long() * nb
        rel_t = rel_t.abs()
    else:
        rel_t = -torch.min(rel_t, torch.zeros_like(rel_t))
    max_exact = nb // 2
    is_small = rel_t < max_exact
    large = max_exact + (torch.log(rel_t.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return (buckets + torch.where(is_small, rel_t, large)).numpy()


def t5_position_bias(table: torch.Tensor, q_len: int, k_len: int, bidirectional=True,
                     num_buckets=32, max_distance=128) -> torch.Tensor:
    """(1,H,q,k) bias from the (num_buckets,H) table — modeling_t5.py:237-251."""
    ctx = np.arange(q_len)[:, None]
    mem = np.arange(k_len)[None, :]
    b = t5_relative_bucket(mem - ctx, bidirectional, num_buckets, max_distance)
    vals = F.embedding(torch.as_tensor(b), table)          # (q,k,H)
    return vals.permute(2, 0, 1).unsqueeze(0)


# -- transformers T5Attention.forward core (modeling_t5.py:312-336), dropout off --------
def t5_attention_core(q, k, v, position_bias, key_valid=None):
    """q,k,v (B,H,S,D).  scores = q k^T (NO 1/sqrt(d)) + bias + mask; fp32 softmax.
    key_valid (B,Sk) bool/float: keys with 0 get finfo.min added (create_bidirectional_mask)."""
    scores = torch.matmul(q, k.transpose(3, 2))
    bias = position_bias
    if key_valid is not None:
        add = torch.zeros(key_valid.shape, dtype=scores.dtype)
        add = add.masked_fill(~key_valid.bool(), torch.finfo(scores.dtype).min)
        bias = bias + add[:, None, None, :]
    scores = scores + bias
    w = F.softmax(scores.float(), dim=-1).type_as(scores)
    return torch.matmul(w, v)


# -- torch nn.MultiheadAttention core as driven by core/model/PhonemeLaTr.py:134-144 ----
def mha_attention_core(q, k, v, causal: bool, key_add=None):
    """q (B,H,T,D), k,v (B,H,S,D).  scores = q k^T / sqrt(D) + causal(-inf) + FLOAT key mask
    added as-is (float masks are additive in nn.MultiheadAttention)."""
    D = q.shape[-1]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(D)
    if causal:
        T, S = q.shape[-2], k.shape[-2]
        m = torch.full((T, S), float("-inf")).triu(1)
        scores = scores + m
    if key_add is not None:
        scores = scores + key_add[:, None, None, :].to(scores.dtype)
    w = F.softmax(scores, dim=-1)
    return torch.matmul(w, v)


# -- core/model/PhonemeLaTr.py:124-130 + core/executor/PhonemeLaTr_Executor.py:181-190 --
def phoneme_head_ce(h, targets, W_on, b_on, W_rh, b_rh, W_to, b_to, ignore_index):
    """h (N,d) = shared_lm_head output; targets (N,3).  Returns (loss, (logits_on, rh, to))."""
    on_dim, rt_dim = W_on.shape[1], W_rh.shape[1]
    lo = F.linear(h[:, :on_dim], W_on, b_on)
    lr = F.linear(h[:, on_dim:on_dim + rt_dim], W_rh, b_rh)
    lt = F.linear(h[:, on_dim + rt_dim:], W_to, b_to)
    loss = (F.cross_entropy(lo, targets[:, 0], ignore_index=ignore_index)
            + F.cross_entropy(lr, targets[:, 1], ignore_index=ignore_index)
            + F.cross_entropy(lt, targets[:, 2], ignore_index=ignore_index))
    return loss, (lo, lr, lt)


# -- core/model/LaTr.py:83 + core/executor/base_executor.py:169 --------------------------
def vocab_head_ce(h, W, targets, ignore_index):
    logits = F.linear(h, W)
    return F.cross_entropy(logits, targets, ignore_index=ignore_index), logits
