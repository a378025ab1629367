"""Autograd-visible wrappers of the C-ABI kernels (host side of the boundary).

Every function here takes CUDA tensors, passes raw device pointers + the current
stream to libpvqa_sm100.so through ctypes, and raises if the library is missing or a
tensor is not on a CUDA device — there is deliberately no PyTorch/CPU fallback
(BASELINE.json north_star: "no CPU fallback, no multi-backend dispatch").
"""
from __future__ import annotations

import os
from ctypes import c_void_p

import torch

from . import _lib
from ._lib import PVQA_BF16, PVQA_F32, check

_DT = {torch.float32: PVQA_F32, torch.bfloat16: PVQA_BF16}


def _dt(t: torch.dtype) -> int:
    try:
        return _DT[t]
    except KeyError:
        raise TypeError(f"libpvqa supports float32 and bfloat16 only, got {t}") from None


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "libpvqa_sm100 ops run on CUDA (sm_100a) tensors only; got a CPU tensor. "
                "There is no CPU fallback in the product path (use oracle/ in tests)."
            )


class KernelTimer:
    """Optional per-kernel CUDA-event timing (bench.py's live roofline numbers).  Events are
    recorded on the launching stream immediately around the C-ABI call."""
    enabled = False
    records: dict = {}

    @classmethod
    def reset(cls, enabled=True):
        cls.enabled = enabled
        cls.records = {}

    @classmethod
    def summary(cls):
        """name -> (launches, total_ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in cls.records.items()}


class _prof:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KernelTimer.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if KernelTimer.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            KernelTimer.records.setdefault(self.name, []).append((self.a, b))
        return False


def _ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


# ----------------------------------------------------------------------------------
# dropout RNG state: (seed, offset) for the counter-based Philox used inside kernels.
# ----------------------------------------------------------------------------------
class _Rng:
    seed = 0x5EED5EED
    offset = 0

    @classmethod
    def next(cls, n_elements: int):
        off = cls.offset
        cls.offset += (n_elements + 7) // 8 + 1
        return cls.seed, off


_RNG_COUNTERS = {}      # device -> int64 tensor, alive as long as the library's pointer to it


def rng_step_counter(device) -> torch.Tensor:
    """The device-resident dropout step counter every dropout kernel adds to its Philox offset (CUDA-graph replays
    bump it to draw fresh masks).  The C side keeps the raw pointer (pvqa_set_rng_step_counter), so the tensor is
    owned HERE, for the life of the process — not by whichever TrainStep happened to create it."""
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    t = _RNG_COUNTERS.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=dev)
        _RNG_COUNTERS[key] = t
    _lib.load().pvqa_set_rng_step_counter(t.data_ptr())
    return t


_ERR_FLAGS = []         # int32 device flags the embedding kernels raise on an out-of-range index


def _new_err_flag(dev) -> torch.Tensor:
    """one persistent flag per device; kernels only ever set it, check_index_errors() reads and clears it"""
    for f in _ERR_FLAGS:
        if f.device == dev:
            return f
    f = torch.zeros(1, dtype=torch.int32, device=dev)
    _ERR_FLAGS.append(f)
    return f


def check_index_errors() -> None:
    """Raise if an embedding kernel saw a token / coordinate / label index outside its table since the last check
    (the reference's nn.Embedding raises a device assert, e.g. for coordinates >= max_2d_position_embeddings; the
    kernels zero the row and set this flag).  One host sync: call it per epoch or every N steps, outside captures."""
    for f in _ERR_FLAGS:
        if int(f.item()) != 0:
            f.zero_()
            raise IndexError("pvqa: an embedding kernel received an index outside its table "
                             "(token id, layout coordinate or phoneme label); the offending rows were zeroed")


def manual_seed(seed: int) -> None:
    """Seed the in-kernel dropout generator (independent from torch's generator)."""
    _Rng.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _Rng.offset = 0


# ----------------------------------------------------------------------------------
# K1: fused multimodal embedding
# ----------------------------------------------------------------------------------
class _EmbedMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_feat, coords, ocr_ids, q_ids, ocr_mask, q_mask, shared_w, out_dtype, *layout_w):
        lib = _lib.load()
        _need_cuda(img_feat, coords, ocr_ids, q_ids, ocr_mask, q_mask, shared_w, *layout_w)
        has_ocr = ocr_ids is not None
        B = q_ids.shape[0] if q_ids is not None else img_feat.shape[0]
        S_img = 0 if img_feat is None else img_feat.shape[1]
        L_ocr = ocr_ids.shape[1] if has_ocr else 0
        L_q = 0 if q_ids is None else q_ids.shape[1]
        V, d = shared_w.shape
        n_pos = layout_w[0].shape[0] if has_ocr else 0
        S = S_img + L_ocr + L_q
        dev = shared_w.device
        if img_feat is not None:
            img_feat = img_feat.to(out_dtype).contiguous()
        if has_ocr:
            assert len(layout_w) == 6, "six layout tables (x0,y0,x1,y1,w,h) are required"
            coords = coords.contiguous()
            ocr_ids = ocr_ids.contiguous()
            ocr_mask = ocr_mask.to(torch.float32).contiguous()
            if coords.dtype != torch.int64 or ocr_ids.dtype != torch.int64:
                raise TypeError("coordinates / tokenized_ocr must be int64")
            layout_w = [w.contiguous() for w in layout_w]
            if any(w.dtype != shared_w.dtype or w.shape != (n_pos, d) for w in layout_w):
                raise TypeError("layout tables must share dtype with the token table and be (n_pos, d)")
        if q_ids is not None:
            q_ids = q_ids.contiguous()
            q_mask = q_mask.to(torch.float32).contiguous()
            if q_ids.dtype != torch.int64:
                raise TypeError("input_ids must be int64")
        shared_c = shared_w.contiguous()
        out = torch.empty((B, S, d), dtype=out_dtype, device=dev)
        out_mask = torch.empty((B, S), dtype=torch.float32, device=dev)
        err = _new_err_flag(dev)      # persistent (no per-call allocation / memset); read by check_index_errors()
        tabs = _ptr_array(layout_w) if has_ocr else None
        with torch.cuda.device(dev), _prof("embed_mm_fwd"):
            check(lib.pvqa_embed_mm_fwd(_p(img_feat), _p(coords), _p(ocr_ids), _p(q_ids), _p(ocr_mask), _p(q_mask),
                                        _p(shared_c), tabs, _p(out), _p(out_mask),
                                        B, S_img, L_ocr, L_q, d, V, n_pos,
                                        _dt(shared_w.dtype), _dt(out_dtype), _p(err), _stream()),
                  "pvqa_embed_mm_fwd")
        ctx.save_for_backward(coords, ocr_ids, q_ids)
        ctx.dims = (B, S_img, L_ocr, L_q, d, V, n_pos)
        ctx.tab_dtype = shared_w.dtype
        ctx.img_needs_grad = img_feat is not None and ctx.needs_input_grad[0]
        ctx.img_dtype = None if img_feat is None else img_feat.dtype
        ctx.mark_non_differentiable(out_mask)
        ctx.err_flag = err
        return out, out_mask

    @staticmethod
    def backward(ctx, d_out, _d_mask):
        lib = _lib.load()
        coords, ocr_ids, q_ids = ctx.saved_tensors
        B, S_img, L_ocr, L_q, d, V, n_pos = ctx.dims
        d_out = d_out.contiguous()
        dev = d_out.device
        d_shared = torch.zeros((V, d), dtype=torch.float32, device=dev)
        d_layout = [torch.zeros((n_pos, d), dtype=torch.float32, device=dev) for _ in range(6)] if L_ocr else []
        tabs = _ptr_array(d_layout) if L_ocr else None
        with torch.cuda.device(dev), _prof("embed_mm_bwd"):
            check(lib.pvqa_embed_mm_bwd(_p(d_out), _p(coords), _p(ocr_ids), _p(q_ids), _p(d_shared), tabs,
                                        B, S_img, L_ocr, L_q, d, V, n_pos, _dt(d_out.dtype), _stream()),
                  "pvqa_embed_mm_bwd")
        d_img = d_out[:, :S_img] if ctx.img_needs_grad else None
        if ctx.tab_dtype != torch.float32:
            d_shared = d_shared.to(ctx.tab_dtype)
            d_layout = [g.to(ctx.tab_dtype) for g in d_layout]
        return (d_img, None, None, None, None, None, d_shared, None, *d_layout)


def embed_multimodal(img_feat, coordinates, tokenized_ocr, input_ids, ocr_attention_mask, src_attention_mask,
                     shared_weight, layout_weights=(), out_dtype=None):
    """K1.  Returns (multi_modal_feat (B,S,d), input_attention_mask (B,S) float32).

    Mirrors `_calculate_embedding` of the reference (core/model/PhonemeLaTr.py:219-231)
    minus the ViT + projector GEMM, whose output is `img_feat`.  `layout_weights` is the
    six SpatialModule tables in coordinate-column order (x0, y0, x1, y1, w, h); pass
    `coordinates=None, tokenized_ocr=None` for the PreSTU family (no layout branch).
    """
    out_dtype = out_dtype or (img_feat.dtype if img_feat is not None else shared_weight.dtype)
    return _EmbedMM.apply(img_feat, coordinates, tokenized_ocr, input_ids, ocr_attention_mask,
                          src_attention_mask, shared_weight, out_dtype, *layout_weights)


def last_index_error(out: torch.Tensor) -> bool:  # pragma: no cover - debugging helper
    fn = out.grad_fn
    flag = getattr(fn, "err_flag", None)
    return bool(flag is not None and int(flag.item()) != 0)


# ----------------------------------------------------------------------------------
# K1': fused phoneme target embedding + sinusoidal PE (+ dropout)
# ----------------------------------------------------------------------------------
class _EmbedTgt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, onset_w, rhyme_w, tone_w, pe, dropout_p, out_dtype):
        lib = _lib.load()
        _need_cuda(labels, onset_w, rhyme_w, tone_w, pe)
        if labels.dtype != torch.int64:
            raise TypeError("labels must be int64 (B,T,3)")
        labels = labels.contiguous()
        B, T, three = labels.shape
        assert three == 3
        V_o, on_dim = onset_w.shape
        V_r, rt_dim = rhyme_w.shape
        V_t, rt2 = tone_w.shape
        d = on_dim + 2 * rt_dim
        if rt2 != rt_dim or pe.shape[-1] != d or pe.shape[-2] < T:
            raise ValueError("phoneme sub-table widths / positional table do not match")
        pe2 = pe.reshape(-1, d)
        if pe2.dtype != torch.float32:
            pe2 = pe2.float()
        pe2 = pe2.contiguous()
        dev = onset_w.device
        out = torch.empty((B, T, d), dtype=out_dtype, device=dev)
        err = _new_err_flag(dev)      # persistent (no per-call allocation / memset); read by check_index_errors()
        seed, offset = _Rng.next(B * T * d) if dropout_p > 0 else (0, 0)
        ow, rw, tw = onset_w.contiguous(), rhyme_w.contiguous(), tone_w.contiguous()
        with torch.cuda.device(dev), _prof("embed_tgt_fwd"):
            check(lib.pvqa_embed_tgt_fwd(_p(labels), _p(ow), _p(rw), _p(tw), _p(pe2), _p(out),
                                         B, T, d, on_dim, rt_dim, V_o, V_r, V_t,
                                         _dt(onset_w.dtype), _dt(out_dtype), float(dropout_p), seed, offset,
                                         _p(err), _stream()),
                  "pvqa_embed_tgt_fwd")
        ctx.save_for_backward(labels)
        ctx.meta = (B, T, d, on_dim, rt_dim, V_o, V_r, V_t, float(dropout_p), seed, offset, onset_w.dtype)
        ctx.err_flag = err
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (labels,) = ctx.saved_tensors
        B, T, d, on_dim, rt_dim, V_o, V_r, V_t, p, seed, offset, tab_dtype = ctx.meta
        d_out = d_out.contiguous()
        dev = d_out.device
        g_on = torch.zeros((V_o, on_dim), dtype=torch.float32, device=dev)
        g_rh = torch.zeros((V_r, rt_dim), dtype=torch.float32, device=dev)
        g_to = torch.zeros((V_t, rt_dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof("embed_tgt_bwd"):
            check(lib.pvqa_embed_tgt_bwd(_p(d_out), _p(labels), _p(g_on), _p(g_rh), _p(g_to),
                                         B, T, d, on_dim, rt_dim, V_o, V_r, V_t, _dt(d_out.dtype),
                                         p, seed, offset, _stream()),
                  "pvqa_embed_tgt_bwd")
        if tab_dtype != torch.float32:
            g_on, g_rh, g_to = g_on.to(tab_dtype), g_rh.to(tab_dtype), g_to.to(tab_dtype)
        return None, g_on, g_rh, g_to, None, None, None


def embed_target(labels, onset_weight, rhyme_weight, tone_weight, pos_embedding, dropout_p=0.0, training=False,
                 out_dtype=None):
    """K1'.  concat(onset[l0], rhyme[l1], tone[l2]) + pos_embedding[:, :T], then dropout.

    Mirrors `positional_encoding(tgt_tok_emb(labels))` (core/model/PhonemeLaTr.py:137-138).
    """
    out_dtype = out_dtype or onset_weight.dtype
    p = float(dropout_p) if training else 0.0
    return _EmbedTgt.apply(labels, onset_weight, rhyme_weight, tone_weight, pos_embedding, p, out_dtype)


# ----------------------------------------------------------------------------------
# Fused glue: RMS norm, residual + dropout, relu + dropout, bf16 linear with fp32 weight gradients
# ----------------------------------------------------------------------------------
class _RmsNorm(torch.autograd.Function):
    """y = rms_norm(x) and, optionally, a pass-through copy of x for the residual path: the backward then
    receives both gradients at once and emits dx = d_residual + d(norm branch) from ONE kernel instead of a
    norm-backward kernel plus an autograd accumulation pass over the fp32 residual stream."""

    @staticmethod
    def forward(ctx, x, weight, eps, out_dtype, with_residual):
        lib = _lib.load()
        _need_cuda(x, weight)
        shape = x.shape
        d = shape[-1]
        x2 = x.reshape(-1, d).contiguous()
        N = x2.shape[0]
        w = weight.to(torch.float32).contiguous()
        y = torch.empty((N, d), dtype=out_dtype, device=x.device)
        rstd = torch.empty(N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device), _prof("rms_norm_fwd"):
            check(lib.pvqa_rms_norm_fwd(_p(x2), _p(w), _p(y), _p(rstd), N, d, float(eps), _dt(x2.dtype), _dt(out_dtype),
                                        _stream()), "pvqa_rms_norm_fwd")
        ctx.save_for_backward(x2, w, rstd)
        ctx.meta = (shape, weight.dtype, with_residual)
        if with_residual:
            return y.view(shape), x.view_as(x)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy, d_res=None):
        lib = _lib.load()
        x2, w, rstd = ctx.saved_tensors
        shape, w_dtype, with_residual = ctx.meta
        N, d = x2.shape
        dy2 = dy.reshape(N, d).contiguous()
        if d_res is not None:
            d_res = d_res.reshape(N, d).to(x2.dtype).contiguous()
        dx = torch.empty_like(x2)
        dw = torch.zeros(d, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device), _prof("rms_norm_bwd"):
            check(lib.pvqa_rms_norm_bwd(_p(dy2), _p(x2), _p(w), _p(rstd), _p(d_res), _p(dx), _p(dw), N, d,
                                        _dt(x2.dtype), _dt(dy2.dtype), _stream()), "pvqa_rms_norm_bwd")
        return dx.view(shape), dw.to(w_dtype), None, None, None


def rms_norm(x, weight, eps, out_dtype):
    """T5LayerNorm (modeling_t5.py:46-70): fp32 variance, no mean subtraction, no bias."""
    if x.dtype == torch.bfloat16:
        out_dtype = torch.bfloat16
    return _RmsNorm.apply(x, weight, eps, out_dtype, False)


def rms_norm_residual(x, weight, eps, out_dtype):
    """(rms_norm(x), x): use the second output for the residual add of a pre-norm block so the two gradient
    paths into x are summed inside the norm-backward kernel."""
    if x.dtype == torch.bfloat16:
        out_dtype = torch.bfloat16
    return _RmsNorm.apply(x, weight, eps, out_dtype, True)


class _ResidualDropoutAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, update, p):
        lib = _lib.load()
        _need_cuda(hidden, update)
        hidden = hidden.contiguous()
        update = update.contiguous()
        n = hidden.numel()
        out = torch.empty_like(hidden)
        seed, off = _Rng.next(n) if p > 0 else (0, 0)
        with torch.cuda.device(hidden.device), _prof("residual_dropout_add"):
            check(lib.pvqa_residual_dropout_add(_p(hidden), _p(update), _p(out), n, _dt(update.dtype), float(p), seed,
                                                off, _stream()), "pvqa_residual_dropout_add")
        ctx.meta = (float(p), seed, off, update.dtype, update.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        p, seed, off, udt, ushape = ctx.meta
        d_out = d_out.contiguous()
        d_upd = torch.empty(ushape, dtype=udt, device=d_out.device)
        with torch.cuda.device(d_out.device), _prof("residual_dropout_bwd"):
            check(lib.pvqa_residual_dropout_bwd(_p(d_out), _p(d_upd), d_out.numel(), _dt(udt), p, seed, off, _stream()),
                  "pvqa_residual_dropout_bwd")
        return d_out, d_upd, None


def residual_dropout_add(hidden, update, p, training):
    """hidden (fp32 residual stream) + dropout(update) in one pass."""
    p = float(p) if training else 0.0
    if hidden.dtype != torch.float32 or hidden.numel() % 8 != 0 or hidden.shape != update.shape:
        raise TypeError("residual_dropout_add expects an fp32 residual stream with numel % 8 == 0")
    return _ResidualDropoutAdd.apply(hidden, update, p)


class _AddDropoutLN(torch.autograd.Function):
    """y = LayerNorm(hidden + dropout(update)); returns (y fp32, y_lp) where y_lp is the bf16 copy the next
    GEMM reads (or None).  One launch forward, one backward (include/pvqa.h: pvqa_add_dropout_ln_*)."""

    @staticmethod
    def forward(ctx, hidden, update, gamma, beta, eps, p, want_lp):
        lib = _lib.load()
        _need_cuda(hidden, update, gamma, beta)
        hidden = hidden.contiguous()
        update = update.contiguous()
        d = hidden.shape[-1]
        N = hidden.numel() // d
        dev = hidden.device
        need_grad = any(ctx.needs_input_grad)
        y = torch.empty_like(hidden)
        y_lp = torch.empty(hidden.shape, dtype=torch.bfloat16, device=dev) if want_lp else None
        z = torch.empty_like(hidden) if need_grad else None
        mean = torch.empty(N, dtype=torch.float32, device=dev) if need_grad else None
        rstd = torch.empty(N, dtype=torch.float32, device=dev) if need_grad else None
        seed, off = _Rng.next(hidden.numel()) if p > 0 else (0, 0)
        with torch.cuda.device(dev), _prof("add_dropout_ln_fwd"):
            check(lib.pvqa_add_dropout_ln_fwd(_p(hidden), PVQA_F32, _p(update), _dt(update.dtype), _p(gamma), _p(beta), _p(z), _p(y),
                                              _p(y_lp), _lib.PVQA_BF16, _p(mean), _p(rstd), N, d, float(eps), float(p),
                                              seed, off, _stream()), "pvqa_add_dropout_ln_fwd")
        ctx.save_for_backward(z, gamma, mean, rstd)
        ctx.meta = (float(p), seed, off, update.dtype, update.shape, N, d)
        ctx.set_materialize_grads(False)
        if want_lp:
            return y, y_lp
        return y, None

    @staticmethod
    def backward(ctx, dy, dy_lp):
        lib = _lib.load()
        z, gamma, mean, rstd = ctx.saved_tensors
        p, seed, off, udt, ushape, N, d = ctx.meta
        dev = z.device
        if dy is None and dy_lp is None:
            return None, None, None, None, None, None, None
        dy = dy.contiguous() if dy is not None else None
        dy_lp = dy_lp.contiguous() if dy_lp is not None else None
        d_hidden = torch.empty_like(z)
        d_upd = torch.empty(ushape, dtype=udt, device=dev)
        dgb = torch.zeros(2, d, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof("add_dropout_ln_bwd"):
            check(lib.pvqa_add_dropout_ln_bwd(_p(dy), _p(dy_lp), _lib.PVQA_BF16, _p(z), _p(gamma), _p(mean), _p(rstd),
                                              _p(d_hidden), _p(d_upd), _dt(udt), _p(dgb[0]), _p(dgb[1]), N, d, p, seed,
                                              off, _stream()), "pvqa_add_dropout_ln_bwd")
        return d_hidden, d_upd, dgb[0], dgb[1], None, None, None


def add_dropout_layer_norm(hidden, update, weight, bias, eps, p, training, want_lp=False):
    """LayerNorm(hidden + dropout(update)) of the post-norm decoder layer; returns (y, y_lp or None)."""
    p = float(p) if training else 0.0
    d = hidden.shape[-1]
    if hidden.dtype != torch.float32 or d % 8 != 0 or d > 1024 or hidden.shape != update.shape:
        raise TypeError("add_dropout_layer_norm expects an fp32 residual stream with d % 8 == 0 and d <= 1024")
    return _AddDropoutLN.apply(hidden, update, weight, bias, float(eps), p, bool(want_lp))


class _AddDropoutRms(torch.autograd.Function):
    """(hidden_out, y) = (hidden + dropout(update), T5LayerNorm(hidden_out)): the residual add that closes one
    T5 sublayer fused with the norm that opens the next (include/pvqa.h: pvqa_add_dropout_rms_*)."""

    @staticmethod
    def forward(ctx, hidden, update, weight, eps, p, out_dtype):
        lib = _lib.load()
        _need_cuda(hidden, update, weight)
        hidden = hidden.contiguous()
        update = update.contiguous()
        d = hidden.shape[-1]
        N = hidden.numel() // d
        dev = hidden.device
        w = weight.to(torch.float32).contiguous()
        hidden_out = torch.empty_like(hidden)
        y = torch.empty(hidden.shape, dtype=out_dtype, device=dev)
        rstd = torch.empty(N, dtype=torch.float32, device=dev)
        seed, off = _Rng.next(hidden.numel()) if p > 0 else (0, 0)
        with torch.cuda.device(dev), _prof("add_dropout_rms_fwd"):
            check(lib.pvqa_add_dropout_rms_fwd(_p(hidden), _p(update), _dt(update.dtype), _p(w), _p(hidden_out), _p(y),
                                               _dt(out_dtype), _p(rstd), N, d, float(eps), float(p), seed, off, _stream()),
                  "pvqa_add_dropout_rms_fwd")
        ctx.save_for_backward(hidden_out, w, rstd)
        ctx.meta = (float(p), seed, off, update.dtype, update.shape, N, d, weight.dtype)
        ctx.set_materialize_grads(False)
        return hidden_out, y

    @staticmethod
    def backward(ctx, d_res, dy):
        lib = _lib.load()
        hidden_out, w, rstd = ctx.saved_tensors
        p, seed, off, udt, ushape, N, d, w_dtype = ctx.meta
        dev = hidden_out.device
        if dy is None:           # the normed output was not used: plain residual add backward
            if d_res is None:
                return None, None, None, None, None, None
            d_res = d_res.contiguous()
            d_upd = torch.empty(ushape, dtype=udt, device=dev)
            with torch.cuda.device(dev), _prof("residual_dropout_bwd"):
                check(lib.pvqa_residual_dropout_bwd(_p(d_res), _p(d_upd), d_res.numel(), _dt(udt), p, seed, off, _stream()),
                      "pvqa_residual_dropout_bwd")
            return d_res, d_upd, None, None, None, None
        dy = dy.contiguous()
        d_res = d_res.to(torch.float32).contiguous() if d_res is not None else None
        d_hidden = torch.empty_like(hidden_out)
        d_upd = torch.empty(ushape, dtype=udt, device=dev)
        dw = torch.zeros(d, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof("add_dropout_rms_bwd"):
            check(lib.pvqa_add_dropout_rms_bwd(_p(dy), _dt(dy.dtype), _p(d_res), _p(hidden_out), _p(w), _p(rstd), _p(d_hidden),
                                               _p(d_upd), _dt(udt), _p(dw), N, d, p, seed, off, _stream()),
                  "pvqa_add_dropout_rms_bwd")
        return d_hidden, d_upd, dw.to(w_dtype), None, None, None


def add_dropout_rms_norm(hidden, update, weight, eps, p, training, out_dtype):
    """-> (hidden + dropout(update), T5LayerNorm of that sum in out_dtype); fp32 residual stream."""
    p = float(p) if training else 0.0
    d = hidden.shape[-1]
    if hidden.dtype != torch.float32 or d % 8 != 0 or d > 1024 or hidden.shape != update.shape:
        raise TypeError("add_dropout_rms_norm expects an fp32 residual stream with d % 8 == 0 and d <= 1024")
    return _AddDropoutRms.apply(hidden, update, weight, float(eps), p, out_dtype)


@torch.no_grad()
def add_layer_norm_lp(hidden, update, weight, bias, eps):
    """Inference-only bf16 stream: x = hidden + update (or hidden itself when update is None), y = LayerNorm(x);
    returns (x, y), both bf16.  The pre-norm step of the frozen ViT tower."""
    lib = _lib.load()
    _need_cuda(hidden, update, weight, bias)
    d = hidden.shape[-1]
    if hidden.dtype != torch.bfloat16 or d % 8 != 0 or d > 1024 or (update is not None and update.dtype != torch.bfloat16):
        raise TypeError("add_layer_norm_lp expects bf16 operands with d % 8 == 0 and d <= 1024")
    hidden = hidden.contiguous()
    N = hidden.numel() // d
    y = torch.empty_like(hidden)
    x = hidden
    if update is not None:
        update = update.contiguous()
        x = torch.empty_like(hidden)
    with torch.cuda.device(hidden.device), _prof("add_ln_lp"):
        check(lib.pvqa_add_dropout_ln_fwd(_p(hidden), PVQA_BF16, _p(update), PVQA_BF16, _p(weight), _p(bias),
                                          _p(x) if update is not None else None, None, _p(y), PVQA_BF16, None, None, N, d,
                                          float(eps), 0.0, 0, 0, _stream()), "pvqa_add_dropout_ln_fwd")
    return x, y


def col_sum(x2d, out=None):
    """fp32 column sums of a (N, d) bf16/fp32 matrix (bias gradients)."""
    lib = _lib.load()
    _need_cuda(x2d)
    x2d = x2d.contiguous()
    N, d = x2d.shape
    if d % 8 != 0:
        return x2d.sum(0, dtype=torch.float32)
    res = torch.empty(d, dtype=torch.float32, device=x2d.device) if out is None else out
    with torch.cuda.device(x2d.device), _prof("col_sum"):
        check(lib.pvqa_col_sum(_p(x2d), _p(res), N, d, _dt(x2d.dtype), 0, _stream()), "pvqa_col_sum")
    return res


class _ReluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        lib = _lib.load()
        _need_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        seed, off = _Rng.next(x.numel()) if p > 0 else (0, 0)
        with torch.cuda.device(x.device), _prof("relu_dropout_fwd"):
            check(lib.pvqa_relu_dropout_fwd(_p(x), _p(y), x.numel(), _dt(x.dtype), float(p), seed, off, _stream()),
                  "pvqa_relu_dropout_fwd")
        ctx.save_for_backward(y)
        ctx.p = float(p)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device), _prof("relu_dropout_bwd"):
            check(lib.pvqa_relu_dropout_bwd(_p(dy), _p(y), _p(dx), y.numel(), _dt(y.dtype), ctx.p, _stream()),
                  "pvqa_relu_dropout_bwd")
        return dx, None


def relu_dropout(x, p, training):
    """dropout(relu(x)) in one pass (T5DenseActDense / TransformerDecoderLayer feed-forward)."""
    if x.numel() % 8 != 0:
        raise TypeError("relu_dropout expects numel % 8 == 0")
    return _ReluDropout.apply(x, float(p) if training else 0.0)


# ----------------------------------------------------------------------------------
# K2 / K3: attention.  bf16 -> tcgen05/TMA flash kernels; fp32 (parity mode) -> fp32 CUDA-core kernels.
# ----------------------------------------------------------------------------------
def _check_attn_operand(t, name, dtype):
    if t.dtype != dtype or t.dim() != 4 or t.stride(3) != 1:
        raise TypeError(f"attention {name} must be a {dtype} (B,S,H,D) view with unit stride on D")


def _st3(t):
    return t.stride(0), t.stride(1), t.stride(2)


def _drop_args(dropout_p, B, H, Sq, Sk):
    if dropout_p <= 0.0:
        return 0.0, 0, 0
    # one Philox counter per (query row, 32-key block) (csrc/attn_fwd.cuh: AttnDrop); _Rng.next reserves n / 8 counters
    seed, off = _Rng.next(8 * B * H * Sq * ((Sk + 31) // 32))
    return float(dropout_p), seed, off


def _scp_args(scp):
    """(bucket u8 (B,L,L), table_t fp32 (H,32), q0) -> ctypes args (bucket, table, q0, L)"""
    if scp is None:
        return None, None, 0, 0
    bucket, table_t, q0 = scp
    return _p(bucket), _p(table_t), int(q0), int(bucket.shape[-1])


def attention_fwd_raw(q, k, v, scale, rel_bias=None, key_add=None, causal=False, drop=(0.0, 0, 0), scp=None):
    """Forward through the C-ABI.  q (B,Sq,H,D), k/v (B,Sk,H,D) strided views (bf16 or fp32).
    Returns (o (B,Sq,H,D) same dtype, lse (B,H,Sq) fp32)."""
    lib = _lib.load()
    _need_cuda(q, k, v, rel_bias, key_add)
    dtype = q.dtype
    if dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("attention supports bfloat16 (tcgen05) and float32 (parity mode)")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _check_attn_operand(t, n, dtype)
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    dev = q.device
    o = torch.empty((B, Sq, H, D), dtype=dtype, device=dev)
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=dev)
    if rel_bias is not None and tuple(rel_bias.shape) != (H, Sq + Sk - 1):
        raise ValueError(f"rel_bias must be (H, Sq+Sk-1) = {(H, Sq + Sk - 1)}, got {tuple(rel_bias.shape)}")
    if key_add is not None and tuple(key_add.shape) != (B, Sk):
        raise ValueError(f"key_add must be (B, Sk) = {(B, Sk)}, got {tuple(key_add.shape)}")
    if dtype != torch.bfloat16:
        fn, name = lib.pvqa_attn_f32_fwd, "attn_f32_fwd"
    else:
        fn, name = lib.pvqa_attn_fwd, "attn_fwd"
    with torch.cuda.device(dev), _prof(f"{name}[Sq={Sq},Sk={Sk}]"):
        check(fn(_p(q), _p(k), _p(v), _p(o), _p(lse), _p(rel_bias), _p(key_add), B, H, Sq, Sk, D,
                 *_st3(q), *_st3(k), *_st3(v), *_st3(o), float(scale), int(bool(causal)),
                 float(drop[0]), int(drop[1]), int(drop[2]), *_scp_args(scp), _stream()), "pvqa_" + name)
    return o, lse


def attention_bwd_raw(q, k, v, o, d_o, lse, scale, rel_bias, key_add, causal, dk, dv, want_d_rel, drop=(0.0, 0, 0),
                      scp=None, want_d_scp=False, rel_far=0):
    """Backward through the C-ABI.  dk/dv are caller-provided (B,Sk,H,D) views (possibly into a packed
    buffer) in q's dtype.  Returns (dq fp32 (B,Sq,H,D), d_rel fp32 or None)."""
    lib = _lib.load()
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    dev = q.device
    d_rel = torch.zeros((H, Sq + Sk - 1), dtype=torch.float32, device=dev) if want_d_rel else None
    d_scp = torch.zeros((H, 32), dtype=torch.float32, device=dev) if (scp is not None and want_d_scp) else None
    sb, st_, sq0, sL = _scp_args(scp)
    d_o = d_o.to(q.dtype)
    if d_o.stride(3) != 1:
        d_o = d_o.contiguous()
    if q.dtype == torch.bfloat16:
        dq = torch.zeros((B, Sq, H, D), dtype=torch.float32, device=dev)      # fp32 accumulator (RED target)
        ws = torch.empty((B, H, Sq), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof(f"attn_bwd[Sq={Sq},Sk={Sk}]"):
            check(lib.pvqa_attn_bwd(_p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(rel_bias), _p(key_add),
                                    _p(dq), _p(dk), _p(dv), _p(d_rel), _p(ws), B, H, Sq, Sk, D,
                                    *_st3(q), *_st3(k), *_st3(v), *_st3(o), *_st3(d_o), *_st3(dk), *_st3(dv),
                                    float(scale), int(bool(causal)), float(drop[0]), int(drop[1]), int(drop[2]),
                                    sb, st_, _p(d_scp), sq0, sL, int(rel_far), _stream()), "pvqa_attn_bwd")
    else:
        dq = torch.empty((B, Sq, H, D), dtype=torch.float32, device=dev)
        ws = torch.empty((B, H, Sq), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof(f"attn_f32_bwd[Sq={Sq},Sk={Sk}]"):
            check(lib.pvqa_attn_f32_bwd(_p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(rel_bias), _p(key_add),
                                        _p(dq), _p(dk), _p(dv), _p(d_rel), _p(ws), B, H, Sq, Sk, D,
                                        *_st3(q), *_st3(k), *_st3(v), *_st3(o), *_st3(d_o), *_st3(dq), *_st3(dk),
                                        *_st3(dv), float(scale), int(bool(causal)), float(drop[0]), int(drop[1]),
                                        int(drop[2]), sb, st_, _p(d_scp), sq0, sL, _stream()), "pvqa_attn_f32_bwd")
    return dq, d_rel, d_scp


def _prep_bias(rel_bias, key_add):
    rb = None if rel_bias is None else rel_bias.detach().to(torch.float32).contiguous()
    ka = None if key_add is None else key_add.detach().to(torch.float32).contiguous()
    return rb, ka


class _AttnSelf(torch.autograd.Function):
    """packed (B,S,3,H,D) projection -> (B,S,H,D); d(qkv) comes back packed for the QKV GEMM backward."""

    @staticmethod
    def forward(ctx, qkv, rel_bias, key_add, scale, causal, dropout_p, scp_bucket, scp_table, scp_q0):
        rb, ka = _prep_bias(rel_bias, key_add)
        qkv = qkv.contiguous()
        B, S, _, H, _ = qkv.shape
        drop = _drop_args(dropout_p, B, H, S, S)
        scp = None
        if scp_bucket is not None:
            if scp_bucket.dtype != torch.uint8 or scp_bucket.dim() != 3:
                raise TypeError("scp_bucket must be uint8 (B,L,L)")
            scp = (scp_bucket.contiguous(), scp_table.detach().to(torch.float32).t().contiguous(), int(scp_q0))
        o, lse = attention_fwd_raw(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale, rb, ka, causal, drop, scp)
        ctx.save_for_backward(qkv, o, lse, rb, ka, *(scp[:2] if scp else ()))
        ctx.meta = (float(scale), bool(causal), rel_bias is not None and ctx.needs_input_grad[1],
                    None if rel_bias is None else rel_bias.dtype, drop,
                    None if scp is None else scp[2], scp is not None and ctx.needs_input_grad[7],
                    None if scp_table is None else scp_table.dtype)
        # producers of bucketed T5 vectors tag them with the start of their constant tails (modules.t5_bucket_far)
        ctx.rel_far = int(getattr(rel_bias, "pvqa_rel_far", 0)) if rel_bias is not None else 0
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse, rb, ka, *scp_t = ctx.saved_tensors
        scale, causal, want_rel, rel_dtype, drop, scp_q0, want_scp, scp_dtype = ctx.meta
        scp = (scp_t[0], scp_t[1], scp_q0) if scp_t else None
        dqkv = torch.empty_like(qkv)
        dq, d_rel, d_scp = attention_bwd_raw(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, d_o, lse, scale, rb, ka,
                                             causal, dqkv[:, :, 1], dqkv[:, :, 2], want_rel, drop, scp, want_scp,
                                             rel_far=ctx.rel_far)
        B_, S_, _, H_, D_ = dqkv.shape
        if (H_ * D_) % 8 == 0:          # fp32 accumulator -> q slot of the packed gradient, one vectorised pass
            with torch.cuda.device(dqkv.device), _prof("cast_rows"):
                check(_lib.load().pvqa_cast_rows(_p(dq), _p(dqkv), B_ * S_, H_ * D_, 3 * H_ * D_, _dt(dqkv.dtype), _stream()),
                      "pvqa_cast_rows")
        else:
            dqkv[:, :, 0].copy_(dq)
        return (dqkv, (d_rel.to(rel_dtype) if want_rel else None), None, None, None, None, None,
                (d_scp.t().to(scp_dtype) if want_scp else None), None)


class _AttnCross(torch.autograd.Function):
    """q (B,Sq,H,D) + packed kv (B,Sk,2,H,D)."""

    @staticmethod
    def forward(ctx, q, kv, rel_bias, key_add, scale, dropout_p):
        rb, ka = _prep_bias(rel_bias, key_add)
        q, kv = q.contiguous(), kv.contiguous()
        B, Sq, H, _ = q.shape
        drop = _drop_args(dropout_p, B, H, Sq, kv.shape[1])
        o, lse = attention_fwd_raw(q, kv[:, :, 0], kv[:, :, 1], scale, rb, ka, False, drop)
        ctx.save_for_backward(q, kv, o, lse, rb, ka)
        ctx.meta = (float(scale), rel_bias is not None and ctx.needs_input_grad[2],
                    None if rel_bias is None else rel_bias.dtype, drop)
        ctx.rel_far = int(getattr(rel_bias, "pvqa_rel_far", 0)) if rel_bias is not None else 0
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, o, lse, rb, ka = ctx.saved_tensors
        scale, want_rel, rel_dtype, drop = ctx.meta
        dkv = torch.empty_like(kv)
        dq, d_rel, _ = attention_bwd_raw(q, kv[:, :, 0], kv[:, :, 1], o, d_o, lse, scale, rb, ka, False,
                                         dkv[:, :, 0], dkv[:, :, 1], want_rel, drop, rel_far=ctx.rel_far)
        return dq.to(q.dtype), dkv, (d_rel.to(rel_dtype) if want_rel else None), None, None, None


def attention_self(qkv, scale, rel_bias=None, key_add=None, causal=False, dropout_p=0.0, scp=None):
    """qkv (B,S,3,H,D) packed projection output (bf16 or fp32).
    scores = scale * q.k + rel_bias[h, j-i+S-1] + key_add[b, j] (+ -inf above the diagonal if causal)
             (+ scp_table[bucket[b, i-q0, j-q0], h] on the OCR block when scp = (bucket u8 (B,L,L), table (32,H), q0))."""
    sb, stab, sq0 = scp if scp is not None else (None, None, 0)
    return _AttnSelf.apply(qkv, rel_bias, key_add, scale, causal, float(dropout_p), sb, stab, sq0)


def attention_cross(q, kv, scale, rel_bias=None, key_add=None, dropout_p=0.0):
    """q (B,Sq,H,D); kv (B,Sk,2,H,D) packed."""
    return _AttnCross.apply(q, kv, rel_bias, key_add, scale, float(dropout_p))


# ----------------------------------------------------------------------------------
# K4: fused phoneme head + 3x cross-entropy (logits never materialised)
# ----------------------------------------------------------------------------------
class _PhonemeHeadCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, targets, W_on, b_on, W_rh, b_rh, W_to, b_to, ignore_index):
        lib = _lib.load()
        _need_cuda(h, targets, W_on, b_on, W_rh, b_rh, W_to, b_to)
        if targets.dtype != torch.int64 or targets.shape[-1] != 3 or targets.stride(-1) != 1:
            raise TypeError("targets must be int64 (N,3) with unit inner stride")
        h = h.contiguous()
        N, d = h.shape
        V_o, on_dim = W_on.shape
        V_r, rt_dim = W_rh.shape
        V_t, _ = W_to.shape
        wdt = h.dtype          # weights are consumed in the activation dtype (fp32 masters cast once: 71k elements)
        ws = [t.to(wdt).contiguous() for t in (W_on, b_on, W_rh, b_rh, W_to, b_to)]
        dev = h.device
        loss_sum = torch.empty(3, dtype=torch.float32, device=dev)
        count = torch.empty(3, dtype=torch.int32, device=dev)
        lse = torch.empty((N, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof("phoneme_head_ce_fwd"):
            check(lib.pvqa_phoneme_head_ce_fwd(_p(h), _p(targets), targets.stride(0), *[_p(w) for w in ws],
                                               _p(loss_sum), _p(count), _p(lse), None, None, None,
                                               N, d, on_dim, rt_dim, V_o, V_r, V_t, int(ignore_index),
                                               _dt(wdt), _dt(h.dtype), _stream()),
                  "pvqa_phoneme_head_ce_fwd")
        # mean over non-ignored targets per head, summed (nan if a head has no valid target, like torch)
        loss = (loss_sum / count.to(torch.float32)).sum()
        ctx.save_for_backward(h, targets, lse, count, *ws)
        ctx.meta = (N, d, on_dim, rt_dim, V_o, V_r, V_t, int(ignore_index), (W_on.dtype, b_on.dtype))
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        h, targets, lse, count, W_on, b_on, W_rh, b_rh, W_to, b_to = ctx.saved_tensors
        N, d, on_dim, rt_dim, V_o, V_r, V_t, ignore_index, (w_dtype, b_dtype) = ctx.meta
        dev = h.device
        g = g.to(torch.float32).reshape(1).contiguous()
        dls = [torch.empty((N, V), dtype=h.dtype, device=dev) for V in (V_o, V_r, V_t)]
        with torch.cuda.device(dev), _prof("phoneme_head_ce_bwd"):
            check(lib.pvqa_phoneme_head_ce_bwd(_p(h), _p(targets), targets.stride(0), _p(W_on), _p(b_on), _p(W_rh),
                                               _p(b_rh), _p(W_to), _p(b_to), _p(lse), _p(count), _p(g),
                                               _p(dls[0]), _p(dls[1]), _p(dls[2]),
                                               N, d, on_dim, rt_dim, V_o, V_r, V_t, ignore_index,
                                               _dt(W_on.dtype), _dt(h.dtype), _stream()),
                  "pvqa_phoneme_head_ce_bwd")
        # the three small GEMMs (cuBLAS): d_h slices, dW_k, db_k
        d_h = torch.empty_like(h)
        offs = (0, on_dim, on_dim + rt_dim)
        widths = (on_dim, rt_dim, rt_dim)
        grads = []
        for dl, W, off, w in zip(dls, (W_on, W_rh, W_to), offs, widths):
            d_h[:, off:off + w] = dl @ W
            grads.append((dl.t() @ h[:, off:off + w]).to(w_dtype))
            grads.append(dl.sum(0, dtype=torch.float32).to(b_dtype))
        return (d_h, None, *grads, None)


class _PhonemeHeadFused(torch.autograd.Function):
    """K4 on tcgen05: loss = 3x CE(heads(shared_lm_head(x))) in one kernel (csrc/head_tc.cu).  The same kernel leaves
    the unscaled logit gradients (softmax - onehot, bf16) behind, so the backward is library GEMMs only: the factor
    g / count_k is folded into the small operands."""

    @staticmethod
    def forward(ctx, x, Ws_lp, Ws, bs, targets, W_on, b_on, W_rh, b_rh, W_to, b_to, ignore_index):
        lib = _lib.load()
        _need_cuda(x, Ws_lp, bs, targets, W_on, b_on, W_rh, b_rh, W_to, b_to)
        if targets.dtype != torch.int64 or targets.shape[-1] != 3 or targets.stride(-1) != 1:
            raise TypeError("targets must be int64 (N,3) with unit inner stride")
        x = x.contiguous()
        N, d = x.shape
        V = (W_on.shape[0], W_rh.shape[0], W_to.shape[0])
        on_dim, rt_dim = W_on.shape[1], W_rh.shape[1]
        dev = x.device
        wk = [t.to(torch.bfloat16).contiguous() for t in (W_on, W_rh, W_to)]
        # biases are consumed at bf16 precision, like the weights (and like the unfused path)
        bk = [t.to(torch.bfloat16).to(torch.float32).contiguous() for t in (b_on, b_rh, b_to)]
        bs32 = bs.to(torch.float32).contiguous()
        h = torch.empty((N, d), dtype=torch.bfloat16, device=dev)
        loss_sum = torch.empty(3, dtype=torch.float32, device=dev)
        count = torch.empty(3, dtype=torch.int32, device=dev)
        lse = torch.empty((N, 3), dtype=torch.float32, device=dev)
        want_grad = any(ctx.needs_input_grad)
        dls = [torch.empty((N, (v + 15) // 16 * 16), dtype=torch.bfloat16, device=dev) for v in V] if want_grad else [None] * 3
        with torch.cuda.device(dev), _prof("phoneme_head_fused_fwd"):
            check(lib.pvqa_phoneme_head_fused_fwd(_p(x), _p(Ws_lp), _p(bs32), _p(targets), targets.stride(0),
                                                  _p(wk[0]), _p(bk[0]), _p(wk[1]), _p(bk[1]), _p(wk[2]), _p(bk[2]),
                                                  _p(h), _p(loss_sum), _p(count), _p(lse), _p(dls[0]), _p(dls[1]), _p(dls[2]),
                                                  N, d, on_dim, rt_dim, V[0], V[1], V[2], int(ignore_index), _stream()),
                  "pvqa_phoneme_head_fused_fwd")
        loss = (loss_sum / count.to(torch.float32)).sum()
        if want_grad:
            ctx.save_for_backward(x, Ws_lp, h, count, *dls, *wk)
        ctx.meta = (V, on_dim, rt_dim, W_on.dtype, b_on.dtype, bs.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, Ws_lp, h, count, dl0, dl1, dl2, w0, w1, w2 = ctx.saved_tensors
        V, on_dim, rt_dim, w_dtype, b_dtype, bs_dtype = ctx.meta
        scale = g.to(torch.float32).reshape(1) / count.to(torch.float32)          # (3,): d loss / d (sum of NLL)_k
        d_h = torch.empty_like(h)
        offs = (0, on_dim, on_dim + rt_dim)
        widths = (on_dim, rt_dim, rt_dim)
        grads = []
        for k, (dl, W, off, w) in enumerate(zip((dl0, dl1, dl2), (w0, w1, w2), offs, widths)):
            # the GEMMs run on the padded (N, round16(V_k)) gradient (pad columns are zero) against zero-padded weight
            # rows: every operand 16-byte aligned (V_t = 7 would send cuBLAS to its `align1` kernels: 62 us for 14 MFLOP)
            hk = h[:, off:off + w]
            Wp = torch.zeros((dl.shape[1], w), dtype=torch.bfloat16, device=dl.device)
            Wp[:V[k]] = W.to(torch.float32) * scale[k]
            torch.mm(dl, Wp, out=d_h[:, off:off + w])
            grads.append((torch.mm(dl.t(), hk, out_dtype=torch.float32)[:V[k]] * scale[k]).to(w_dtype))
            grads.append((dl.sum(0, dtype=torch.float32)[:V[k]] * scale[k]).to(b_dtype))
        dx = d_h @ Ws_lp if ctx.needs_input_grad[0] else None
        dWs = torch.mm(d_h.t(), x, out_dtype=torch.float32)
        dbs = col_sum(d_h).to(bs_dtype)
        return (dx, None, dWs, dbs, None, *grads, None)


def phoneme_head_fused(x, W_shared_lp, W_shared, b_shared, targets, W_onset, b_onset, W_rhyme, b_rhyme, W_tone, b_tone,
                       ignore_index):
    """K4, one tcgen05 kernel: x (N,768) bf16 = decoder output; W_shared_lp = bf16 shadow of the fp32 master W_shared
    (the gradient goes to the master).  Returns the scalar onset+rhyme+tone cross-entropy."""
    return _PhonemeHeadFused.apply(x, W_shared_lp, W_shared, b_shared, targets, W_onset, b_onset, W_rhyme, b_rhyme,
                                   W_tone, b_tone, ignore_index)


def phoneme_head_fused_supported(x, W_onset, W_rhyme, W_tone):
    return (x.dtype == torch.bfloat16 and x.shape[-1] == 768 and W_onset.shape[1] == 256 and W_rhyme.shape[1] == 256
            and max(W_onset.shape[0], W_rhyme.shape[0], W_tone.shape[0]) <= 192)


def phoneme_head_ce(h, targets, W_onset, b_onset, W_rhyme, b_rhyme, W_tone, b_tone, ignore_index):
    """K4.  h (N,d) = shared_lm_head output, targets (N,3) int64.  Returns the scalar
    onset+rhyme+tone cross-entropy of core/executor/PhonemeLaTr_Executor.py:181-190."""
    return _PhonemeHeadCE.apply(h, targets, W_onset, b_onset, W_rhyme, b_rhyme, W_tone, b_tone, ignore_index)


# ----------------------------------------------------------------------------------
# K4 (large vocabulary, LaTr): chunked lm_head + cross-entropy; full logits never exist
# ----------------------------------------------------------------------------------
class _VocabHeadCE(torch.autograd.Function):
    """loss = CE(h @ W^T, targets).  Forward walks the rows in chunks: cuBLAS logits chunk (fp32) ->
    pvqa_vocab_ce_grad (loss + bf16 dlogits in one pass) -> cuBLAS dh chunk and dW accumulation.  The gradients
    are therefore produced during forward (like fused linear-cross-entropy implementations) and only scaled
    by the incoming grad in backward."""

    @staticmethod
    def forward(ctx, h, weight, w_lp, targets, ignore_index, chunk_rows):
        lib = _lib.load()
        _need_cuda(h, weight, targets)
        N, d = h.shape
        V = weight.shape[0]
        dev = h.device
        targets = targets.reshape(-1).contiguous()
        cdt = h.dtype
        w_c = w_lp if w_lp is not None else weight.to(cdt)
        count = (targets != ignore_index).sum().to(torch.float32)
        inv_count = (1.0 / count).reshape(1)          # inf -> nan loss when nothing is valid, like torch
        loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        need_h = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1]
        d_h = torch.empty((N, d), dtype=cdt, device=dev) if need_h else None
        d_w = torch.zeros((V, d), dtype=torch.float32, device=dev) if need_w else None
        hb = h if cdt == torch.bfloat16 else None
        for r0 in range(0, N, chunk_rows):
            r1 = min(N, r0 + chunk_rows)
            hc = h[r0:r1]
            logits = torch.mm(hc, w_c.t(), out_dtype=torch.float32) if cdt == torch.bfloat16 else hc @ w_c.t()
            dl = torch.empty((r1 - r0, V), dtype=torch.bfloat16, device=dev)
            with torch.cuda.device(dev), _prof("vocab_ce_grad"):
                check(lib.pvqa_vocab_ce_grad(_p(logits), _p(targets[r0:r1]), 1, _p(inv_count), _p(loss_sum), _p(dl),
                                             r1 - r0, V, int(ignore_index), _stream()), "pvqa_vocab_ce_grad")
            if cdt == torch.bfloat16:
                if need_h:
                    torch.mm(dl, w_c, out=d_h[r0:r1])
                if need_w:
                    d_w = torch.addmm(d_w, dl.t(), hc, out_dtype=torch.float32)
            else:                      # fp32 parity mode: keep the GEMMs in fp32
                dlf = dl.float()
                if need_h:
                    torch.mm(dlf, w_c, out=d_h[r0:r1])
                if need_w:
                    d_w.addmm_(dlf.t(), hc)
            del logits, dl
        ctx.save_for_backward(d_h, d_w)
        ctx.w_dtype = weight.dtype
        return (loss_sum * inv_count).reshape(())

    @staticmethod
    def backward(ctx, g):
        d_h, d_w = ctx.saved_tensors
        gh = None if d_h is None else d_h * g.to(d_h.dtype)
        gw = None if d_w is None else (d_w * g).to(ctx.w_dtype)
        return gh, gw, None, None, None, None


def vocab_head_ce(h, weight, targets, ignore_index, w_lp=None, chunk_rows=1024):
    """K4-large.  h (N,d) decoder output, weight (V,d) lm_head (fp32 master; w_lp = optional bf16 shadow),
    targets (N,) int64.  Returns mean cross-entropy over non-ignored targets."""
    return _VocabHeadCE.apply(h.contiguous(), weight, w_lp, targets, ignore_index, int(chunk_rows))
