"""Host-side building blocks with the reference's module / parameter names.

Each class mirrors a module of the reference (or of the libraries it calls) closely enough
that ``state_dict()`` keys and tensor shapes are identical, so checkpoints written by the
reference load with ``strict=True``.  Dense linears stay cuBLAS (torch);
the gather/scatter embeddings, attention and the phoneme head + loss call libpvqa_sm100.so
through ``ops``.

Precision: parameters are fp32 masters.  ``compute_dtype`` (fp32 or bf16) selects the
activation dtype of linears / attention / kernels; the residual stream and all
normalisation statistics stay fp32, which is what ``torch.autocast(bfloat16)`` gives the
reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _ShadowWeights:
    """Low-precision shadows of the fp32 master weights used by the bf16 compute path.

    A shadow may be the row-wise concatenation of several masters (packed q|k|v projection).
    Masters are compared by `_version` (bumped by the optimizer's in-place update); when any
    entry is stale ALL registered shadows are refreshed with one multi-tensor copy, so a training
    step pays one cast launch instead of one per linear per direction."""

    def __init__(self):
        self.entries = {}     # key -> dict(masters, shadow, views, versions)

    def get(self, masters, dtype):
        key = tuple(id(m) for m in masters) + (dtype,)
        e = self.entries.get(key)
        if e is None or e["shadow"].device != masters[0].device or any(r() is not m for r, m in zip(e["refs"], masters)):
            import weakref
            rows = sum(m.shape[0] for m in masters)
            shadow = torch.empty((rows,) + tuple(masters[0].shape[1:]), dtype=dtype, device=masters[0].device)
            views, off = [], 0
            for m in masters:
                views.append(shadow[off:off + m.shape[0]])
                off += m.shape[0]
            e = {"refs": [weakref.ref(m) for m in masters], "shadow": shadow, "views": views, "versions": None}
            self.entries[key] = e
        if e["versions"] != tuple(m._version for m in masters):
            self.refresh()
        return e["shadow"]

    def invalidate(self):
        """force the next lookup to refresh (used right before CUDA-graph capture)"""
        for e in self.entries.values():
            e["versions"] = None

    @torch.no_grad()
    def refresh(self):
        dst, src = [], []
        dead = []
        for key, e in self.entries.items():
            ms = [r() for r in e["refs"]]
            if any(m is None for m in ms):
                dead.append(key)
                continue
            ver = tuple(m._version for m in ms)
            if e["versions"] != ver:
                dst.extend(e["views"])
                src.extend(m.detach() for m in ms)
                e["versions"] = ver
        for key in dead:
            del self.entries[key]
        if dst:
            torch._foreach_copy_(dst, src)


SHADOWS = _ShadowWeights()


class _LinearLP(torch.autograd.Function):
    """y = x W^T + b with W, b given as low-precision shadows; weight/bias gradients are produced
    directly in fp32 (bf16 x bf16 -> fp32 GEMM output) and routed to the fp32 masters."""

    @staticmethod
    def forward(ctx, x, w_lp, b_lp, n_masters, *masters):
        ctx.save_for_backward(x, w_lp)
        ctx.n_w = n_masters
        ctx.has_bias = b_lp is not None
        ctx.rows = [m.shape[0] for m in masters[:n_masters]]
        import weakref
        ctx.slot_of = weakref.ref(masters[0]) if n_masters == 1 else None
        return F.linear(x, w_lp, b_lp)

    @staticmethod
    def backward(ctx, dy):
        x, w_lp = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = (dy2 @ w_lp).view(x.shape) if ctx.needs_input_grad[0] else None
        # data-parallel runs: a single-master weight gradient is written straight into its all-reduce bucket slot
        # (parallel.grad_slot) and that view is returned as the gradient — autograd adopts it, the reducer's hook sees
        # its own storage and skips the copy
        slot = None
        if ctx.n_w == 1 and ctx.slot_of is not None:
            from .parallel import grad_slot
            slot = grad_slot(ctx.slot_of())
        if slot is not None:
            dW = torch.mm(dy2.t(), x2, out_dtype=torch.float32, out=slot)
        else:
            dW = torch.mm(dy2.t(), x2, out_dtype=torch.float32)
        grads, off = [], 0
        for r in ctx.rows:
            grads.append(dW[off:off + r] if len(ctx.rows) > 1 else dW)
            off += r
        if ctx.has_bias:
            db = ops.col_sum(dy2)
            off = 0
            for r in ctx.rows:
                grads.append(db[off:off + r])
                off += r
        return (dx, None, None, None, *grads)


def _lin_multi(x, weights, biases=None):
    """Linear with the row-wise concatenation of `weights` (and `biases`); fp32 masters."""
    if x.dtype == weights[0].dtype:
        w = weights[0] if len(weights) == 1 else torch.cat(list(weights), dim=0)
        b = None if biases is None else (biases[0] if len(biases) == 1 else torch.cat(list(biases), dim=0))
        return F.linear(x, w, b)
    w_lp = SHADOWS.get(list(weights), x.dtype)
    b_lp = None if biases is None else SHADOWS.get(list(biases), x.dtype)
    masters = list(weights) + (list(biases) if biases is not None else [])
    return _LinearLP.apply(x, w_lp, b_lp, len(weights), *masters)


def _lin(x, weight, bias=None):
    """F.linear in x's dtype with fp32 master weights."""
    return _lin_multi(x, [weight], None if bias is None else [bias])


def _lin_rows(x, weight, bias, r0, r1):
    """Linear with rows [r0, r1) of a packed master weight/bias (nn.MultiheadAttention in_proj slices)."""
    if x.dtype == weight.dtype:
        return F.linear(x, weight[r0:r1], bias[r0:r1])
    w_lp = SHADOWS.get([weight], x.dtype)[r0:r1]
    b_lp = SHADOWS.get([bias], x.dtype)[r0:r1]
    return _LinearLP.apply(x, w_lp, b_lp, 1, weight[r0:r1], bias[r0:r1])


# ----------------------------------------------------------------------------------
# reference: core/model/PhonemeLaTr.py:17-44 (identical copies in LaTr.py, CustomizedLaTr.py)
# ----------------------------------------------------------------------------------
class SpatialModule(nn.Module):
    """Six 2-D layout tables.  forward() is only used stand-alone; inside the models the
    six gathers are fused with the token embedding and the concat by K1."""

    ORDER = ("top_left_x", "top_left_y", "bottom_right_x", "bottom_right_y", "width_emb", "height_emb")

    def __init__(self, config):
        super().__init__()
        n, d = config.max_2d_position_embeddings, config.d_model
        # registration order follows the reference so state_dict() key order matches too
        self.top_left_x = nn.Embedding(n, d)
        self.bottom_right_x = nn.Embedding(n, d)
        self.top_left_y = nn.Embedding(n, d)
        self.bottom_right_y = nn.Embedding(n, d)
        self.width_emb = nn.Embedding(n, d)
        self.height_emb = nn.Embedding(n, d)

    def tables(self):
        """weights in coordinate-column order x0, y0, x1, y1, w, h"""
        return [getattr(self, n).weight for n in self.ORDER]

    def forward(self, coordinates):
        B, L, _ = coordinates.shape
        w = self.top_left_x.weight
        zero_tok = torch.zeros((1, w.shape[1]), dtype=w.dtype, device=w.device)
        ids = torch.zeros((B, L), dtype=torch.long, device=coordinates.device)
        ones = torch.ones((B, L), dtype=torch.float32, device=coordinates.device)
        out, _ = ops.embed_multimodal(None, coordinates, ids, None, ones, None, zero_tok, self.tables(),
                                      out_dtype=w.dtype)
        return out


# ----------------------------------------------------------------------------------
# reference: PhonoLaTr/modules.py:27-63 (3-table form expected by core/model/PhonemeLaTr.py:72-78)
# ----------------------------------------------------------------------------------
class PhonemeEmbedding(nn.Module):
    def __init__(self, onset_vocab_size, rhyme_vocab_size, tone_vocab_size, onset_embed_dim, rhyme_tone_embed_dim):
        super().__init__()
        self.onset_embedding = nn.Embedding(onset_vocab_size, onset_embed_dim)
        self.rhyme_embedding = nn.Embedding(rhyme_vocab_size, rhyme_tone_embed_dim)
        self.tone_embedding = nn.Embedding(tone_vocab_size, rhyme_tone_embed_dim)

    def forward(self, phoneme_tensor, pos_embedding=None, dropout_p=0.0, training=False, out_dtype=None):
        d = self.onset_embedding.embedding_dim + 2 * self.rhyme_embedding.embedding_dim
        if pos_embedding is None:
            pos_embedding = torch.zeros((phoneme_tensor.shape[1], d), device=phoneme_tensor.device)
        return ops.embed_target(phoneme_tensor, self.onset_embedding.weight, self.rhyme_embedding.weight,
                                self.tone_embedding.weight, pos_embedding, dropout_p=dropout_p,
                                training=training, out_dtype=out_dtype)


# reference: core/model/modules/transformer_utils.py:27-36
class TokenEmbedding(nn.Module):
    """flat target vocabulary (char / byte / BPE ids): embedding * sqrt(d)"""

    def __init__(self, vocab_size: int, emb_size):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, emb_size)
        self.emb_size = emb_size

    def forward(self, tokens):
        return F.embedding(tokens.long(), self.embedding.weight) * math.sqrt(self.emb_size)


# reference: core/model/modules/transformer_utils.py:6-25
class SinusoidalPositionalEncoding(nn.Module):
    def __init__(self, emb_size: int, dropout: float, maxlen: int = 5000):
        super().__init__()
        den = torch.exp(-torch.arange(0, emb_size, 2) * math.log(10000) / emb_size)
        pos = torch.arange(0, maxlen).reshape(maxlen, 1)
        pos_embedding = torch.zeros((maxlen, emb_size))
        pos_embedding[:, 0::2] = torch.sin(pos * den)
        pos_embedding[:, 1::2] = torch.cos(pos * den)
        self.p = dropout
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("pos_embedding", pos_embedding.unsqueeze(0))

    def forward(self, token_embedding):
        return self.dropout(token_embedding + self.pos_embedding[:, : token_embedding.size(1)].to(token_embedding.dtype))


# ----------------------------------------------------------------------------------
# T5 encoder (HF transformers T5Stack layout: modeling_t5.py:46-70,153-500,617-793)
# ----------------------------------------------------------------------------------
class T5LayerNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x, out_dtype=None):
        return ops.rms_norm(x, self.weight, self.variance_epsilon, out_dtype or x.dtype)

    def with_residual(self, x, out_dtype=None):
        """-> (normed, x) with the residual gradient fused into the norm backward"""
        return ops.rms_norm_residual(x, self.weight, self.variance_epsilon, out_dtype or x.dtype)


_BUCKET_CACHE: dict = {}


def t5_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128):
    """HF T5Attention._relative_position_bucket (modeling_t5.py:190-235) on a LongTensor, evaluated with the
    same torch float32 ops so bucket ids are bit-exact."""
    nb = num_buckets
    buckets = torch.zeros_like(rel)
    if bidirectional:
        nb //= 2
        buckets = buckets + (rel > 0).long() * nb
        rp = rel.abs()
    else:
        rp = -torch.min(rel, torch.zeros_like(rel))
    max_exact = nb // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rp, large)


def t5_bucket_lut(q_len, k_len, bidirectional, num_buckets, max_distance, device):
    """bucket id for every relative offset rel = j - i in [-(q_len-1), k_len-1] (HF formula,
    modeling_t5.py:190-235, evaluated with the same torch float32 ops so ids are bit-exact)."""
    key = (q_len, k_len, bidirectional, num_buckets, max_distance, str(device))
    lut = _BUCKET_CACHE.get(key)
    if lut is None:
        rel = torch.arange(-(q_len - 1), k_len, dtype=torch.long)
        nb = num_buckets
        buckets = torch.zeros_like(rel)
        if bidirectional:
            nb //= 2
            buckets = buckets + (rel > 0).long() * nb
            rp = rel.abs()
        else:
            rp = -torch.min(rel, torch.zeros_like(rel))
        max_exact = nb // 2
        is_small = rp < max_exact
        large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact)
                             * (nb - max_exact)).long()
        large = torch.min(large, torch.full_like(large, nb - 1))
        lut = (buckets + torch.where(is_small, rp, large)).to(device)
        _BUCKET_CACHE[key] = lut
    return lut


def t5_bucket_far(q_len, k_len, bidirectional, num_buckets, max_distance):
    """Smallest F >= 1 such that the bucket id is one constant for every offset rel >= F and one constant for every
    rel <= -F (with 32 buckets / max distance 128: F = 91).  The attention backward uses it to credit the bias gradient
    of blocks that lie entirely in those tails to a single offset (include/pvqa.h: rel_far) — exact, because the table
    gradient sums d_rel over all offsets of a bucket anyway.  0 when the vector has no such tails."""
    key = ("far", q_len, k_len, bidirectional, num_buckets, max_distance)
    far = _BUCKET_CACHE.get(key)
    if far is None:
        lut = t5_bucket_lut(q_len, k_len, bidirectional, num_buckets, max_distance, "cpu")
        zero = q_len - 1                                   # index of rel = 0
        far = 0
        for f in range(1, min(q_len, k_len)):
            pos, neg = lut[zero + f:], lut[:zero - f + 1]
            if bool((pos == pos[0]).all()) and bool((neg == neg[0]).all()):
                far = f
                break
        _BUCKET_CACHE[key] = far
    return far


class T5Attention(nn.Module):
    def __init__(self, config, has_relative_attention_bias=False):
        super().__init__()
        self.is_decoder = getattr(config, "is_decoder", False)
        self.has_relative_attention_bias = has_relative_attention_bias
        self.relative_attention_num_buckets = config.relative_attention_num_buckets
        self.relative_attention_max_distance = config.relative_attention_max_distance
        self.d_model = config.d_model
        self.key_value_proj_dim = config.d_kv
        self.n_heads = config.num_heads
        self.dropout = config.dropout_rate
        self.inner_dim = self.n_heads * self.key_value_proj_dim
        self.q = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.k = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.v = nn.Linear(self.d_model, self.inner_dim, bias=False)
        self.o = nn.Linear(self.inner_dim, self.d_model, bias=False)
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(self.relative_attention_num_buckets, self.n_heads)

    def rel_bias(self, q_len, k_len):
        """(H, q_len + k_len - 1) fp32: bias as a function of the relative offset j - i."""
        lut = t5_bucket_lut(q_len, k_len, not self.is_decoder, self.relative_attention_num_buckets,
                            self.relative_attention_max_distance, self.relative_attention_bias.weight.device)
        rb = self.relative_attention_bias.weight.float()[lut].t().contiguous()
        rb.pvqa_rel_far = t5_bucket_far(q_len, k_len, not self.is_decoder, self.relative_attention_num_buckets,
                                        self.relative_attention_max_distance)
        return rb

    def forward(self, x, rel_bias, key_add, kv=None, causal=False, scp=None):
        """x (B,S,d) compute dtype.  Self-attention when kv is None, else cross-attention on kv."""
        B, S, _ = x.shape
        H, D = self.n_heads, self.key_value_proj_dim
        p_drop = self.dropout if self.training else 0.0
        if kv is None:
            qkv = _lin_multi(x, [self.q.weight, self.k.weight, self.v.weight]).view(B, S, 3, H, D)
            o = ops.attention_self(qkv, scale=1.0, rel_bias=rel_bias, key_add=key_add, causal=causal,
                                   dropout_p=p_drop, scp=scp)
        else:
            q = _lin(x, self.q.weight).view(B, S, H, D)
            kvp = _lin_multi(kv, [self.k.weight, self.v.weight]).view(B, kv.shape[1], 2, H, D)
            o = ops.attention_cross(q, kvp, scale=1.0, rel_bias=rel_bias, key_add=key_add, dropout_p=p_drop)
        return _lin(o.reshape(B, S, H * D), self.o.weight)


class T5LayerSelfAttention(nn.Module):
    def __init__(self, config, has_relative_attention_bias=False):
        super().__init__()
        self.SelfAttention = T5Attention(config, has_relative_attention_bias)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = nn.Dropout(config.dropout_rate)

    def core(self, normed, rel_bias, key_add, causal=False, scp=None):
        """the sublayer between its norm and its residual add (T5Stack fuses those two with the neighbours)"""
        return self.SelfAttention(normed, rel_bias, key_add, causal=causal, scp=scp)

    def forward(self, hidden, rel_bias, key_add, compute_dtype, causal=False, scp=None):
        normed, hidden = self.layer_norm.with_residual(hidden, out_dtype=compute_dtype)
        attn = self.core(normed, rel_bias, key_add, causal=causal, scp=scp)
        return ops.residual_dropout_add(hidden, attn, self.dropout.p, self.training)


class T5LayerCrossAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.EncDecAttention = T5Attention(config, has_relative_attention_bias=False)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = nn.Dropout(config.dropout_rate)

    def core(self, normed, memory, key_add):
        return self.EncDecAttention(normed, None, key_add, kv=memory)

    def forward(self, hidden, memory, key_add, compute_dtype):
        normed, hidden = self.layer_norm.with_residual(hidden, out_dtype=compute_dtype)
        attn = self.core(normed, memory, key_add)
        return ops.residual_dropout_add(hidden, attn, self.dropout.p, self.training)


class T5DenseActDense(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.wi = nn.Linear(config.d_model, config.d_ff, bias=False)
        self.wo = nn.Linear(config.d_ff, config.d_model, bias=False)
        self.dropout = nn.Dropout(config.dropout_rate)
        self.act_name = config.dense_act_fn
        if self.act_name != "relu":
            # VietAI/vit5-base / -large (every YAML of the reference) are ReLU T5s; there is no PyTorch fall-through:
            # unsupported shapes are errors
            raise NotImplementedError(f"T5 feed-forward activation {self.act_name!r}: the B200 path implements the "
                                      "ReLU feed-forward of the reference's backbones only")

    def forward(self, x):
        h = _lin(x, self.wi.weight)
        h = ops.relu_dropout(h, self.dropout.p, self.training)
        return _lin(h, self.wo.weight)


class T5LayerFF(nn.Module):
    def __init__(self, config):
        super().__init__()
        if getattr(config, "is_gated_act", False):
            raise NotImplementedError("gated T5 feed-forward (T5 v1.1): not used by the reference's backbones, not "
                                      "implemented on the B200 path (no PyTorch fall-through)")
        self.DenseReluDense = T5DenseActDense(config)
        self.layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = nn.Dropout(config.dropout_rate)

    def core(self, normed):
        return self.DenseReluDense(normed)

    def forward(self, hidden, compute_dtype):
        normed, hidden = self.layer_norm.with_residual(hidden, out_dtype=compute_dtype)
        ff = self.core(normed)
        return ops.residual_dropout_add(hidden, ff, self.dropout.p, self.training)


class T5Block(nn.Module):
    def __init__(self, config, has_relative_attention_bias=False, is_decoder=False):
        super().__init__()
        self.is_decoder = is_decoder
        self.layer = nn.ModuleList([T5LayerSelfAttention(config, has_relative_attention_bias)])
        if is_decoder:
            self.layer.append(T5LayerCrossAttention(config))
        self.layer.append(T5LayerFF(config))

    def forward(self, hidden, rel_bias, key_add, compute_dtype, memory=None, memory_key_add=None, causal=False,
                scp=None):
        hidden = self.layer[0](hidden, rel_bias, key_add, compute_dtype, causal=causal, scp=scp)
        if self.is_decoder:
            hidden = self.layer[1](hidden, memory, memory_key_add, compute_dtype)
        return self.layer[-1](hidden, compute_dtype)


def _t5_init(module, config):
    """HF T5PreTrainedModel._init_weights (factor = config.initializer_factor)."""
    factor = config.initializer_factor
    d_model, d_kv, n_heads, d_ff = config.d_model, config.d_kv, config.num_heads, config.d_ff
    for m in module.modules():
        if isinstance(m, T5LayerNorm):
            nn.init.constant_(m.weight, factor * 1.0)
        elif isinstance(m, T5DenseActDense):
            nn.init.normal_(m.wi.weight, mean=0.0, std=factor * (d_model ** -0.5))
            nn.init.normal_(m.wo.weight, mean=0.0, std=factor * (d_ff ** -0.5))
        elif isinstance(m, T5Attention):
            nn.init.normal_(m.q.weight, mean=0.0, std=factor * ((d_model * d_kv) ** -0.5))
            nn.init.normal_(m.k.weight, mean=0.0, std=factor * (d_model ** -0.5))
            nn.init.normal_(m.v.weight, mean=0.0, std=factor * (d_model ** -0.5))
            nn.init.normal_(m.o.weight, mean=0.0, std=factor * ((n_heads * d_kv) ** -0.5))
            if m.has_relative_attention_bias:
                nn.init.normal_(m.relative_attention_bias.weight, mean=0.0, std=factor * (d_model ** -0.5))


class T5Stack(nn.Module):
    """Encoder (or decoder) stack with HF's parameter names: embed_tokens, block.{i}.layer.{j}..., final_layer_norm."""

    def __init__(self, config, embed_tokens=None, is_decoder=False, num_layers=None):
        super().__init__()
        self.config = config
        self.is_decoder = is_decoder
        self.embed_tokens = embed_tokens if embed_tokens is not None else nn.Embedding(config.vocab_size, config.d_model)
        n = num_layers if num_layers is not None else config.num_layers
        self.block = nn.ModuleList([T5Block(config, has_relative_attention_bias=(i == 0), is_decoder=is_decoder)
                                    for i in range(n)])
        for b in self.block:
            b.layer[0].SelfAttention.is_decoder = is_decoder
        self.final_layer_norm = T5LayerNorm(config.d_model, eps=config.layer_norm_epsilon)
        self.dropout = nn.Dropout(config.dropout_rate)

    @staticmethod
    def key_add_from_mask(attention_mask):
        """HF create_bidirectional_mask semantics (masking_utils, probed in SURVEY §8c): key j is
        masked iff attention_mask[b, j] == 0; masked keys get finfo.min (== -inf for softmax)."""
        if attention_mask is None:
            return None
        return torch.where(attention_mask != 0, 0.0, float("-inf")).to(torch.float32)

    def forward(self, inputs_embeds, attention_mask=None, compute_dtype=torch.float32, memory=None,
                memory_mask=None, external_rel_bias=None, scp=None):
        hidden = F.dropout(inputs_embeds.float(), self.dropout.p, self.training)
        S = hidden.shape[1]
        if external_rel_bias is not None:
            rel_bias = external_rel_bias     # SaL: bias supplied by the caller, mask NOT added
            key_add = None
        else:
            rel_bias = self.block[0].layer[0].SelfAttention.rel_bias(S, S)
            key_add = self.key_add_from_mask(attention_mask)
        mem = None if memory is None else memory.to(compute_dtype)
        mem_key_add = self.key_add_from_mask(memory_mask) if memory is not None else None
        d = hidden.shape[-1]
        if d % 8 != 0 or d > 1024:
            raise ValueError(f"d_model = {d}: the fused norm kernels keep a row in registers (d % 8 == 0, d <= 1024); "
                             "there is no PyTorch fall-through")
        if len(self.block) == 0:
            hidden = self.final_layer_norm(hidden, out_dtype=torch.float32)
            return F.dropout(hidden, self.dropout.p, self.training)
        # Pre-norm chain: every `hidden + dropout(sublayer)` is fused with the T5LayerNorm that opens the NEXT
        # sublayer (the last one with final_layer_norm): one launch instead of two per sublayer, forward and backward.
        subs = []
        for blk in self.block:
            sa = blk.layer[0]
            subs.append((sa, lambda n, sa=sa: sa.core(n, rel_bias, key_add, causal=self.is_decoder, scp=scp)))
            if blk.is_decoder:
                ca = blk.layer[1]
                subs.append((ca, lambda n, ca=ca: ca.core(n, mem, mem_key_add)))
            subs.append((blk.layer[-1], blk.layer[-1].core))
        normed, hidden = subs[0][0].layer_norm.with_residual(hidden, out_dtype=compute_dtype)
        for i, (sub, core) in enumerate(subs):
            last = i + 1 == len(subs)
            nrm = self.final_layer_norm if last else subs[i + 1][0].layer_norm
            hidden, normed = ops.add_dropout_rms_norm(hidden, core(normed), nrm.weight, nrm.variance_epsilon,
                                                      sub.dropout.p, self.training,
                                                      torch.float32 if last else compute_dtype)
        return F.dropout(normed, self.dropout.p, self.training)


class T5ForConditionalGeneration(nn.Module):
    """`.shared`, `.encoder`, `.decoder`, `.lm_head` with HF's names and weight tying
    (transformers T5ForConditionalGeneration; used by the plain LaTr / PreSTU family, core/model/LaTr.py:47)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = T5Stack(config, self.shared, is_decoder=False)
        n_dec = getattr(config, "num_decoder_layers", None) or config.num_layers
        self.decoder = T5Stack(config, self.shared, is_decoder=True, num_layers=n_dec)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        _t5_init(self, config)
        nn.init.normal_(self.shared.weight, mean=0.0, std=config.initializer_factor * 1.0)
        if getattr(config, "tie_word_embeddings", True):
            self.lm_head.weight = self.shared.weight
        else:
            nn.init.normal_(self.lm_head.weight, mean=0.0, std=config.initializer_factor * 1.0)


class T5EncoderModel(nn.Module):
    """`.shared` + `.encoder` exactly like HF's T5EncoderModel (tied embed_tokens)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = T5Stack(config, self.shared, is_decoder=False)
        _t5_init(self, config)
        nn.init.normal_(self.shared.weight, mean=0.0, std=config.initializer_factor * 1.0)


# ----------------------------------------------------------------------------------
# target-side decoder: torch nn.TransformerDecoder parameter layout
# reference: core/model/modules/transformer_utils.py:38-64
# ----------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    """Parameter holder with nn.MultiheadAttention's names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, d_model, n_head):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, n_head
        self.head_dim = d_model // n_head
        assert self.head_dim * n_head == d_model, "embed_dim must be divisible by num_heads"
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class TransformerDecoderLayer(nn.Module):
    """Post-norm decoder layer (torch defaults: relu, dim_feedforward 2048, dropout .1, eps 1e-5)."""

    def __init__(self, d_model, n_head, dim_feedforward=2048, dropout=0.1, layer_norm_eps=1e-5):
        super().__init__()
        self.self_attn = _MHAParams(d_model, n_head)
        self.multihead_attn = _MHAParams(d_model, n_head)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm3 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.p = dropout

    def _tail(self, x, update, norm, compute_dtype):
        """x = norm(x + dropout(update)) in one launch; also returns the compute-dtype copy the next GEMM reads."""
        y, y_lp = ops.add_dropout_layer_norm(x, update, norm.weight, norm.bias, norm.eps, self.p, self.training,
                                             want_lp=(compute_dtype == torch.bfloat16))
        return y, (y_lp if y_lp is not None else y.to(compute_dtype))

    def forward(self, x, memory, tgt_key_add, memory_key_add, compute_dtype, causal=True, xc=None):
        """-> (x fp32 residual stream, xc = x in compute_dtype for the next layer's first GEMM)"""
        B, T, d = x.shape
        sa, ca = self.self_attn, self.multihead_attn
        H, D = sa.num_heads, sa.head_dim
        p_attn = self.p if self.training else 0.0
        scale = 1.0 / math.sqrt(D)
        if xc is None:
            xc = x.to(compute_dtype)
        qkv = _lin(xc, sa.in_proj_weight, sa.in_proj_bias).view(B, T, 3, H, D)
        a = ops.attention_self(qkv, scale=scale, rel_bias=None, key_add=tgt_key_add, causal=causal, dropout_p=p_attn)
        a = _lin(a.reshape(B, T, d), sa.out_proj.weight, sa.out_proj.bias)
        x, xc = self._tail(x, a, self.norm1, compute_dtype)
        q = _lin_rows(xc, ca.in_proj_weight, ca.in_proj_bias, 0, d).view(B, T, H, D)
        kv = _lin_rows(memory, ca.in_proj_weight, ca.in_proj_bias, d, 3 * d).view(B, memory.shape[1], 2, H, D)
        c = ops.attention_cross(q, kv, scale=scale, rel_bias=None, key_add=memory_key_add, dropout_p=p_attn)
        c = _lin(c.reshape(B, T, d), ca.out_proj.weight, ca.out_proj.bias)
        x, xc = self._tail(x, c, self.norm2, compute_dtype)
        h = ops.relu_dropout(_lin(xc, self.linear1.weight, self.linear1.bias), self.p, self.training)
        h = _lin(h, self.linear2.weight, self.linear2.bias)
        return self._tail(x, h, self.norm3, compute_dtype)


class DecoderCache:
    """Per-layer key/value cache for incremental (one new token per call) decoding: the cross-attention K/V of
    the encoder memory are projected once, the self-attention K/V grow by one position per step.  The
    reference re-runs the whole 4-layer decoder over the growing prefix every step (core/model/PhonemeLaTr.py
    :193-215, O(T^2) layer passes); with the cache a step is O(1) layer passes and reads only the prefix K/V."""

    def __init__(self, layers, memory, max_len, compute_dtype):
        B, S, d = memory.shape
        l0 = layers[0]
        H, D = l0.self_attn.num_heads, l0.self_attn.head_dim
        self.layers = layers
        self.len = 0
        self.max_len = max_len
        self.self_kv = [torch.empty((B, max_len, 2, H, D), dtype=compute_dtype, device=memory.device) for _ in layers]
        self.mem_kv = [torch.empty((B, S, 2, H, D), dtype=compute_dtype, device=memory.device) for _ in layers]
        self.set_memory(memory)

    @torch.no_grad()
    def reorder(self, src):
        """beam search: row n of the self-attention cache continues row src[n] (the memory K/V are per sample and
        identical across the beams of a sample, so they stay where they are)"""
        for li, buf in enumerate(self.self_kv):
            self.self_kv[li] = buf.index_select(0, src)

    @torch.no_grad()
    def set_memory(self, memory):
        """project a new encoder memory into the (address-stable) cross-attention K/V buffers and rewind"""
        B, S, d = memory.shape
        for buf, layer in zip(self.mem_kv, self.layers):
            ca = layer.multihead_attn
            buf.copy_(_lin_rows(memory.to(buf.dtype), ca.in_proj_weight, ca.in_proj_bias, d, 3 * d).view(buf.shape))
        self.len = 0


class _DecoderStack(nn.Module):
    def __init__(self, d_model, n_head, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([TransformerDecoderLayer(d_model, n_head) for _ in range(num_layers)])
        # nn.TransformerDecoder deep-copies ONE initialised layer: all layers start identical
        for layer in self.layers[1:]:
            layer.load_state_dict(self.layers[0].state_dict())
        self.num_layers = num_layers


class BaseDecoder(nn.Module):
    """reference: core/model/modules/transformer_utils.py:38-64.  The float masks are ADDED
    to the scores exactly like nn.MultiheadAttention does with float masks:
    tgt_key_padding_mask (1.0 = pad) and memory_key_padding_mask (1.0 = valid) are per-key
    additive terms, the square-subsequent mask is the kernel's causal flag."""

    def __init__(self, emb_size: int, num_layers: int, n_head: int, batch_first: bool = True):
        super().__init__()
        assert batch_first
        self.decoder = _DecoderStack(emb_size, n_head, num_layers)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, compute_dtype=torch.float32, causal=True):
        if memory_mask is not None:
            raise NotImplementedError("memory_mask is never passed by the reference models")

        def as_add(m):
            if m is None:
                return None
            if m.dtype == torch.bool:          # bool masks mask (PhonemeSaL passes bool)
                return torch.where(m, float("-inf"), 0.0).to(torch.float32)
            return m.to(torch.float32)         # float masks are additive

        tka, mka = as_add(tgt_key_padding_mask), as_add(memory_key_padding_mask)
        x = tgt.float()
        mem = memory.to(compute_dtype)
        xc = None
        for layer in self.decoder.layers:
            x, xc = layer(x, mem, tka, mka, compute_dtype, causal=causal, xc=xc)
        return x

    # ---- incremental decoding (inference only) ----
    def new_cache(self, memory, max_len, compute_dtype):
        return DecoderCache(self.decoder.layers, memory.to(compute_dtype), max_len, compute_dtype)

    @torch.no_grad()
    def step(self, x_t, cache, memory_key_padding_mask=None, compute_dtype=torch.float32):
        """x_t (B,1,d): embedding (+PE) of the newest target position.  Returns the decoder output for that
        position — identical (up to float reassociation) to row t of the full causal pass."""
        mka = None if memory_key_padding_mask is None else memory_key_padding_mask.to(torch.float32).contiguous()
        x = x_t.float()
        B, _, d = x.shape
        t = cache.len
        assert t < cache.max_len, "decoder cache is full"
        if d % 8 != 0 or d > 1024:
            raise ValueError(f"d_model = {d}: the fused norm kernels need d % 8 == 0 and d <= 1024")
        lp = compute_dtype == torch.bfloat16

        def tail(x, upd, norm):
            """norm(x + upd) -> (fp32 stream, compute-dtype copy): the same fused launch the full layer pass uses"""
            y, y_lp = ops.add_dropout_layer_norm(x, upd, norm.weight, norm.bias, norm.eps, 0.0, False, want_lp=lp)
            return y, (y_lp if lp else y)

        xc = x.to(compute_dtype)
        for li, layer in enumerate(self.decoder.layers):
            sa, ca = layer.self_attn, layer.multihead_attn
            H, D = sa.num_heads, sa.head_dim
            scale = 1.0 / math.sqrt(D)
            qkv = _lin(xc, sa.in_proj_weight, sa.in_proj_bias).view(B, 1, 3, H, D)
            cache.self_kv[li][:, t] = qkv[:, 0, 1:]
            kv = cache.self_kv[li][:, : t + 1]
            a, _ = ops.attention_fwd_raw(qkv[:, :, 0].contiguous(), kv[:, :, 0], kv[:, :, 1], scale)
            a = _lin(a.reshape(B, 1, d), sa.out_proj.weight, sa.out_proj.bias)
            x, xc = tail(x, a, layer.norm1)
            q = _lin_rows(xc, ca.in_proj_weight, ca.in_proj_bias, 0, d).view(B, 1, H, D)
            mkv = cache.mem_kv[li]
            c, _ = ops.attention_fwd_raw(q, mkv[:, :, 0], mkv[:, :, 1], scale, None, mka)
            c = _lin(c.reshape(B, 1, d), ca.out_proj.weight, ca.out_proj.bias)
            x, xc = tail(x, c, layer.norm2)
            h = torch.relu(_lin(xc, layer.linear1.weight, layer.linear1.bias))
            h = _lin(h, layer.linear2.weight, layer.linear2.bias)
            x, xc = tail(x, h, layer.norm3)
        cache.len = t + 1
        return x


# ----------------------------------------------------------------------------------
# SaL family: 1-D + spatial (SCP) relative position bias
# reference: core/model/modules/SaL_utils.py:24-223.  The reference materialises (B,H,S,S) fp32 and makes a
# GPU -> CPU -> numpy -> GPU round trip per forward (:161-168); here the 1-D part is the (H, 2S-1) relative
# vector of the attention kernels and the SCP part is a uint8 bucket map (B, L_ocr, L_ocr) computed on the
# device from a 121 x 121 cell-distance LUT, looked up inside the kernels.
# ----------------------------------------------------------------------------------
class _RelBiasTable(nn.Module):
    def __init__(self, num_heads, num_buckets=32):
        super().__init__()
        self.relative_attention_bias = nn.Embedding(num_buckets, num_heads)


class RelativePositionBias1D(_RelBiasTable):
    def rel_vector(self, S):
        lut = t5_bucket_lut(S, S, True, 32, 128, self.relative_attention_bias.weight.device)
        rb = self.relative_attention_bias.weight.float()[lut].t().contiguous()
        rb.pvqa_rel_far = t5_bucket_far(S, S, True, 32, 128)
        return rb


class SCPRelativePositionBias(_RelBiasTable):
    GRID = 11

    def __init__(self, num_heads, num_buckets=32):
        super().__init__(num_heads, num_buckets)
        g = self.GRID
        xs, ys = np.mgrid[0:g, 0:g]
        cells = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=1).astype(np.float64)       # cell id = x * 11 + y
        dist = np.sqrt(((cells[:, None, :] - cells[None, :, :]) ** 2).sum(-1)) * 5           # SaL_utils.py:171-195
        rel = torch.tensor(dist).to(torch.long)                                               # truncation, :166
        lut = t5_bucket(rel, bidirectional=True, num_buckets=num_buckets, max_distance=100)   # :66-73 with :127
        self.register_buffer("cell_bucket_lut", lut.to(torch.uint8), persistent=False)

    def buckets(self, coordinates):
        """coordinates (B,L,4) in [0,1) -> uint8 (B,L,L) bucket ids (SaL_utils.py:152-168, on the device)."""
        g = self.GRID
        # mean of two values == (a + b) / 2 exactly; list indexing would build a host index tensor (not capturable)
        xc = (coordinates[:, :, 0] + coordinates[:, :, 2]) / 2
        yc = (coordinates[:, :, 1] + coordinates[:, :, 3]) / 2
        cx = torch.floor(xc * g).to(torch.long).clamp_(0, g - 1)
        cy = torch.floor(yc * g).to(torch.long).clamp_(0, g - 1)
        cell = cx * g + cy
        return self.cell_bucket_lut[cell[:, :, None], cell[:, None, :]].contiguous()


class RelativePositionBiasAggregated(nn.Module):
    def __init__(self, Relative1D, SCP):
        super().__init__()
        self.Relative1D = Relative1D
        self.SCP = SCP

    def forward(self, S, coordinates, max_ques, max_ocr):
        """-> (rel_vector (H, 2S-1), (scp buckets u8 (B,L,L), scp table (32,H), q0))"""
        return self.Relative1D.rel_vector(S), (self.SCP.buckets(coordinates[:, :max_ocr]),
                                               self.SCP.relative_attention_bias.weight, int(max_ques))

    def dense(self, S, coordinates, max_ques, max_ocr):
        """the reference's materialised (B,H,S,S) tensor (tests / debugging only)"""
        rel, (bk, tab, q0) = self.forward(S, coordinates, max_ques, max_ocr)
        i = torch.arange(S, device=rel.device)
        dense = rel[:, (i[None, :] - i[:, None] + S - 1)][None].repeat(coordinates.shape[0], 1, 1, 1)
        L = bk.shape[-1]
        dense[:, :, q0:q0 + L, q0:q0 + L] += tab.float()[bk.long()].permute(0, 3, 1, 2)
        return dense
