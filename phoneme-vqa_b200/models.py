"""Drop-in model classes: same names, constructor / forward / generate signatures and
state_dict keys as the reference's ``core.model`` classes, with the hot
path (fused embeddings, attention, phoneme head + loss) running in libpvqa_sm100.so.

Selecting them from the reference's YAML is ``MODEL_CLASS: "PhonemeLaTr"`` after
``from phoneme_vqa_b200.models import *`` has been added to ``core/model/__init__.py``
(see INTEGRATION.md).
"""
from __future__ import annotations

import contextlib
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .modules import (SHADOWS, BaseDecoder, RelativePositionBias1D, RelativePositionBiasAggregated,
                      SCPRelativePositionBias, T5LayerNorm, PhonemeEmbedding, SinusoidalPositionalEncoding, SpatialModule, T5EncoderModel,
                      T5ForConditionalGeneration, TokenEmbedding, _lin)

__all__ = ["LaTr_config", "PreSTU_config", "SaL_config", "CustomizedLaTr_config", "CustomizedPreSTU_config",
           "CustomizedSaL_config", "LaTr", "PreSTU", "SaL", "CustomizedLaTr", "CustomizedPreSTU", "CustomizedSaL",
           "PhonemeLaTr", "PhonemePreSTU", "PhonemeSaL", "phoneme_beam_search", "reference_beam_select"]


def _random_init(config) -> bool:
    return os.environ.get("PVQA_RANDOM_INIT", "0") == "1" or bool(getattr(config, "random_init", False))


def _phoneme_head_loss(m, dec, targets, ignore_index):
    """K4: shared_lm_head + the three heads + 3x cross-entropy.  bf16 at d = 768: ONE tcgen05 kernel (csrc/head_tc.cu);
    other dtypes / widths: library GEMM for shared_lm_head + the mma.sync / fp32 head kernel."""
    x = dec.reshape(-1, dec.shape[-1])
    tg = targets.reshape(-1, 3)
    heads = (m.onset_lm_head.weight, m.onset_lm_head.bias, m.rhyme_lm_head.weight, m.rhyme_lm_head.bias,
             m.tone_lm_head.weight, m.tone_lm_head.bias)
    if ops.phoneme_head_fused_supported(x, heads[0], heads[2], heads[4]):
        from .modules import SHADOWS
        w_lp = SHADOWS.get([m.shared_lm_head.weight], torch.bfloat16)
        return ops.phoneme_head_fused(x, w_lp, m.shared_lm_head.weight, m.shared_lm_head.bias, tg, *heads, ignore_index)
    h = _lin(dec, m.shared_lm_head.weight, m.shared_lm_head.bias)
    return ops.phoneme_head_ce(h.reshape(-1, h.shape[-1]), tg, *heads, ignore_index)


def _auto_config(name):
    from transformers import AutoConfig
    return AutoConfig.from_pretrained(name)


# reference: core/model/LaTr.py:5-12
class LaTr_config:
    def build(self, config):
        model_config = _auto_config(config.backbone_name)
        model_config.update({"max_2d_position_embeddings": config.max_2d_position_embeddings,
                             "vit_model": config.vit_model_name})
        return model_config


# reference: core/model/PreSTU.py:5-11
class PreSTU_config:
    def build(self, config):
        model_config = _auto_config(config.backbone_name)
        model_config.update({"vit_model": config.vit_model_name})
        return model_config


# reference: core/model/SaL.py:13-21
class SaL_config:
    def build(self, config, new_token_embedding_size):
        model_config = _auto_config(config.backbone_name)
        model_config.update({"ocr_hidden": config.ocr_hidden, "obj_hidden": config.obj_hidden,
                             "new_token_embedding_size": new_token_embedding_size})
        return model_config


# reference: core/model/PhonemeLaTr.py:6-15 (same class in CustomizedLaTr.py)
class CustomizedLaTr_config:
    def build(self, config):
        model_config = _auto_config(config.encoder_name)
        model_config.update({"max_2d_position_embeddings": config.max_2d_position_embeddings,
                             "vit_model": config.vit_model_name,
                             "num_decoder_layers": config.num_decoder_layers,
                             "n_head": config.n_head})
        return model_config


# reference: core/model/CustomizedPreSTU.py:6-14 (backbone_name) / core/model/PhonemePreSTU.py:6-14 (encoder_name);
# every YAML that selects these models carries both keys with the same value
class CustomizedPreSTU_config:
    def build(self, config):
        model_config = _auto_config(getattr(config, "backbone_name", None) or config.encoder_name)
        model_config.update({"vit_model": config.vit_model_name,
                             "num_decoder_layers": config.num_decoder_layers,
                             "n_head": config.n_head})
        return model_config


def _build_vit(config):
    """HF ViTModel (library code: the ViT tower is outside the three north-star kernels).
    Pretrained weights like the reference unless random init was requested (no network here)."""
    from transformers import ViTConfig, ViTModel
    if _random_init(config):
        vc = getattr(config, "vit_config", None)
        vcfg = ViTConfig(**vc) if isinstance(vc, dict) else ViTConfig()
        vcfg._attn_implementation = "sdpa"
        return ViTModel(vcfg)
    return ViTModel.from_pretrained(config.vit_model, attn_implementation="sdpa")


def _load_pretrained_t5_encoder(encoder: T5EncoderModel, config):
    if _random_init(config):
        return
    from transformers import T5EncoderModel as HFEncoder
    sd = HFEncoder.from_pretrained(config._name_or_path).state_dict()
    encoder.load_state_dict(sd, strict=True)


class _FrozenVitLP:
    """bf16 forward of the frozen ViT tower (HF ViTModel semantics: pre-norm blocks, exact GELU, final layernorm;
    transformers models/vit/modeling_vit.py, run by the reference at core/model/PhonemeLaTr.py:220).  Weights are
    bf16 copies of the fp32 `vit.*` parameters (q/k/v fused into one projection), LayerNorm parameters stay fp32.
    Per block: 4 GEMMs, the tcgen05 attention kernel on the packed (B,S,3,H,D) projection (no transposes or
    copies), and two fused `x += sublayer; y = LayerNorm(x)` launches."""

    def __init__(self, vit):
        import copy
        cfg = vit.config
        self.eps = float(cfg.layer_norm_eps)
        self.heads = int(cfg.num_attention_heads)
        self.lean = (cfg.hidden_size // self.heads == 64 and cfg.hidden_size % 8 == 0 and cfg.hidden_size <= 1024
                     and cfg.hidden_act == "gelu")
        bf = torch.bfloat16
        if not self.lean:                        # unusual geometry: the HF modules in bf16
            self.full = copy.deepcopy(vit).to(bf).eval()
            for p in self.full.parameters():
                p.requires_grad_(False)
            return
        self.embeddings = copy.deepcopy(vit.embeddings).to(bf).eval()     # only for unexpected image sizes
        for p in self.embeddings.parameters():
            p.requires_grad_(False)
        # patch embedding as unfold + GEMM (the stride-16 16x16 convolution is exactly that; cuDNN's conv path
        # spends 10x longer in layout conversions), [CLS] and the position table folded in
        pe = vit.embeddings.patch_embeddings
        self.patch = tuple(pe.patch_size) if hasattr(pe.patch_size, "__len__") else (pe.patch_size, pe.patch_size)
        self.image = tuple(pe.image_size) if hasattr(pe.image_size, "__len__") else (pe.image_size, pe.image_size)
        w = pe.projection.weight.detach()
        self.w_patch = w.reshape(w.shape[0], -1).to(bf).contiguous()
        self.b_patch = pe.projection.bias.detach().to(bf)
        pos = vit.embeddings.position_embeddings.detach()
        self.cls_pos = (vit.embeddings.cls_token.detach() + pos[:, :1]).to(bf)
        self.pos = pos[:, 1:].to(bf).contiguous()
        self.layers = []
        for blk in vit.encoder.layer:
            att = blk.attention.attention
            self.layers.append(dict(
                ln1=(blk.layernorm_before.weight.detach().float(), blk.layernorm_before.bias.detach().float()),
                ln2=(blk.layernorm_after.weight.detach().float(), blk.layernorm_after.bias.detach().float()),
                wqkv=torch.cat([att.query.weight, att.key.weight, att.value.weight]).detach().to(bf),
                bqkv=torch.cat([att.query.bias, att.key.bias, att.value.bias]).detach().to(bf),
                wo=blk.attention.output.dense.weight.detach().to(bf), bo=blk.attention.output.dense.bias.detach().to(bf),
                w1=blk.intermediate.dense.weight.detach().to(bf), b1=blk.intermediate.dense.bias.detach().to(bf),
                w2=blk.output.dense.weight.detach().to(bf), b2=blk.output.dense.bias.detach().to(bf)))
        self.final_ln = (vit.layernorm.weight.detach().float(), vit.layernorm.bias.detach().float())

    def _embed(self, px):
        """HF ViTEmbeddings.forward (patch projection, [CLS], position table; dropout is 0 in eval) -> (B, 1+N, d)"""
        bf = torch.bfloat16
        B, C, Hh, Ww = px.shape
        (ph, pw), d = self.patch, self.w_patch.shape[0]
        if (Hh, Ww) != self.image or C * ph * pw != self.w_patch.shape[1]:
            return self.embeddings(px.to(bf)).contiguous()
        gh, gw = Hh // ph, Ww // pw
        patches = torch.empty((B, gh, gw, C, ph, pw), dtype=bf, device=px.device)
        patches.copy_(px.view(B, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5))       # unfold + cast, one pass
        x = torch.empty((B, 1 + gh * gw, d), dtype=bf, device=px.device)
        x[:, :1] = self.cls_pos
        torch.add(F.linear(patches.view(B, gh * gw, C * ph * pw), self.w_patch, self.b_patch), self.pos, out=x[:, 1:])
        return x

    @torch.no_grad()
    def __call__(self, pixel_values):
        bf = torch.bfloat16
        if not self.lean:
            emb = self.full.embeddings(pixel_values.to(bf))
            return self.full.layernorm(self.full.encoder(emb).last_hidden_state)
        x = self._embed(pixel_values)
        B, S, d = x.shape
        H, D = self.heads, d // self.heads
        scale = 1.0 / math.sqrt(D)
        n = len(self.layers)
        _, y = ops.add_layer_norm_lp(x, None, *self.layers[0]["ln1"], self.eps) if n else (x, x)
        for i, L in enumerate(self.layers):
            qkv = F.linear(y, L["wqkv"], L["bqkv"]).view(B, S, 3, H, D)
            a, _ = ops.attention_fwd_raw(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)
            x, y = ops.add_layer_norm_lp(x, F.linear(a.view(B, S, d), L["wo"], L["bo"]), *L["ln2"], self.eps)
            m = F.linear(F.gelu(F.linear(y, L["w1"], L["b1"])), L["w2"], L["b2"])
            nxt = self.layers[i + 1]["ln1"] if i + 1 < n else self.final_ln
            x, y = ops.add_layer_norm_lp(x, m, *nxt, self.eps)
        if n == 0:
            _, y = ops.add_layer_norm_lp(x, None, *self.final_ln, self.eps)
        return y


class _VisionMixin:
    """frozen-ViT handling shared by the LaTr / PreSTU families."""

    def _vit_frozen(self):
        return not any(p.requires_grad for p in self.vit.parameters())

    def _vit_shadow(self):
        """bf16 copy of the FROZEN ViT tower used for bf16 compute.  It is deliberately kept out of the module
        tree (so `state_dict()` still holds exactly the reference's fp32 `vit.*` tensors) and is rebuilt when
        the fp32 weights change (load_state_dict, .to(device))."""
        sig = tuple((p.data_ptr(), p._version) for p in self.vit.parameters())
        cache = self.__dict__.get("_vit_lp")
        if cache is None or cache[0] != sig:
            cache = (sig, _FrozenVitLP(self.vit))
            self.__dict__["_vit_lp"] = cache
        return cache[1]

    def _vit_tokens(self, pixel_values):
        # ViTModel.forward minus the pooler (never used by the reference: PhonemeLaTr.py:220)
        if self._vit_frozen() and self.compute_dtype == torch.bfloat16 and pixel_values.is_cuda:
            return self._vit_shadow()(pixel_values)
        ctx = torch.no_grad() if self._vit_frozen() else contextlib.nullcontext()
        amp = (torch.autocast("cuda", dtype=torch.bfloat16) if self.compute_dtype == torch.bfloat16
               else contextlib.nullcontext())
        with ctx, amp:
            emb = self.vit.embeddings(pixel_values)
            seq = self.vit.encoder(emb).last_hidden_state
            seq = self.vit.layernorm(seq)
        return seq

    def set_compute_dtype(self, dtype):
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        return self


class PhonemeLaTr(nn.Module, _VisionMixin):
    """reference: core/model/PhonemeLaTr.py:46-236"""

    def __init__(self, config, onset_vocab_size, rhyme_vocab_size, tone_vocab_size):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)

        self.spatial_feat_extractor = SpatialModule(config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)

        # freeze ViT (reference :63-66)
        for _, child in self.vit.named_children():
            for param in child.parameters():
                param.requires_grad = False

        d = config.d_model
        self.rhyme_tone_embed_dim = d // 3
        self.onset_embed_dim = int(d - self.rhyme_tone_embed_dim * 2)
        self.tgt_tok_emb = PhonemeEmbedding(onset_vocab_size, rhyme_vocab_size, tone_vocab_size,
                                            self.onset_embed_dim, self.rhyme_tone_embed_dim)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(emb_size=d, num_layers=config.num_decoder_layers, n_head=config.n_head)
        self.shared_lm_head = nn.Linear(d, d)
        self.onset_lm_head = nn.Linear(self.onset_embed_dim, onset_vocab_size)
        self.rhyme_lm_head = nn.Linear(self.rhyme_tone_embed_dim, rhyme_vocab_size)
        self.tone_lm_head = nn.Linear(self.rhyme_tone_embed_dim, tone_vocab_size)

    # -- reference :219-231 ---------------------------------------------------------
    def _calculate_embedding(self, pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask,
                             tokenized_ocr):
        vit_tokens = self._vit_tokens(pixel_values)
        img_feat = _lin(vit_tokens.to(self.compute_dtype), self.visual_projector.weight, self.visual_projector.bias)
        return ops.embed_multimodal(img_feat, coordinates, tokenized_ocr, input_ids, ocr_attention_mask,
                                    src_attention_mask, self.encoder.shared.weight,
                                    self.spatial_feat_extractor.tables(), out_dtype=self.compute_dtype)

    def _encode(self, pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask, tokenized_ocr):
        inputs_embeds, attention_mask = self._calculate_embedding(
            pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask, tokenized_ocr)
        enc = self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        return enc, attention_mask

    # -- reference :134-144 ---------------------------------------------------------
    def decode(self, labels, encoder_outputs, encoder_attention_mask, label_attention_mask=None):
        emb = self.tgt_tok_emb(labels, self.positional_encoding.pos_embedding,
                               dropout_p=self.positional_encoding.p, training=self.training,
                               out_dtype=torch.float32)
        # square-subsequent float mask == causal flag; float key masks are additive
        return self.decoder(emb, encoder_outputs, tgt_mask=None,
                            memory_key_padding_mask=encoder_attention_mask,
                            tgt_key_padding_mask=label_attention_mask,
                            compute_dtype=self.compute_dtype, causal=True)

    def _heads(self, h):
        on, rt = self.onset_embed_dim, self.rhyme_tone_embed_dim
        return (_lin(h[:, :, :on], self.onset_lm_head.weight, self.onset_lm_head.bias),
                _lin(h[:, :, on:on + rt], self.rhyme_lm_head.weight, self.rhyme_lm_head.bias),
                _lin(h[:, :, on + rt:], self.tone_lm_head.weight, self.tone_lm_head.bias))

    # -- reference :98-132 ----------------------------------------------------------
    def forward(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                ocr_attention_mask, tokenized_ocr):
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        dec = self.decode(labels, enc, attention_mask, label_attention_mask)
        h = _lin(dec.to(self.compute_dtype), self.shared_lm_head.weight, self.shared_lm_head.bias)
        on, rh, to = self._heads(h)
        return on.float(), rh.float(), to.float()

    def forward_loss(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                     ocr_attention_mask, tokenized_ocr, targets, ignore_index):
        """Fused fast path: model forward + the executor's 3x CrossEntropyLoss
        (core/executor/PhonemeLaTr_Executor.py:181-190) without materialising logits.
        `targets` = labels[:, 1:, :] of the executor, `labels` = labels[:, :-1]."""
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        dec = self.decode(labels, enc, attention_mask, label_attention_mask)
        return _phoneme_head_loss(self, dec.to(self.compute_dtype), targets, ignore_index)

    # -- reference :146-217 ---------------------------------------------------------
    def generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask, tokenized_ocr,
                 start_symbol, end_symbol, max_length=20, isgreedy=True, num_beam=2, use_graph=False):
        return self.greedy_generate(pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                                    tokenized_ocr, start_symbol, end_symbol, max_length, use_graph=use_graph)

    @torch.no_grad()
    def greedy_generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                        tokenized_ocr, start_symbol, end_symbol, max_len=100, use_cache=True, use_graph=False):
        """reference :169-217.  `use_cache=True` decodes incrementally with a key/value cache;
        `use_cache=False` re-runs the decoder over the growing prefix exactly like the reference loop;
        `use_graph=True` (with the cache) replays one captured CUDA graph per target position — a decode step is
        ~100 small launches, i.e. launch-bound when issued from Python."""
        bz = input_ids.size(0)
        dev = input_ids.device
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        if use_graph and use_cache and dev.type == "cuda":
            return self._greedy_graphed(enc, attention_mask, start_symbol, end_symbol, max_len)
        ys = torch.tensor([[[start_symbol, 0, 0]]], dtype=torch.long, device=dev).repeat(bz, 1, 1)
        cache = self.decoder.new_cache(enc, max_len + 1, self.compute_dtype) if use_cache else None
        pe = self.positional_encoding.pos_embedding
        for t in range(max_len):
            if use_cache:
                emb = self.tgt_tok_emb(ys[:, -1:], pe[:, t:t + 1], out_dtype=torch.float32)
                out = self.decoder.step(emb, cache, attention_mask, self.compute_dtype)
            else:
                out = self.decode(ys, enc, attention_mask)[:, -1:]
            # NOTE: like the reference (:195-205) greedy decoding does NOT apply shared_lm_head
            on, rh, to = self._heads(out.to(self.compute_dtype))
            nxt = torch.stack([on[:, -1].float().argmax(-1), rh[:, -1].float().argmax(-1),
                               to[:, -1].float().argmax(-1)], dim=-1)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
            if torch.any(ys[:, :, 0] == end_symbol, dim=1).sum() == bz:
                break
        return ys

    # -- beam search over the factorised (onset, rhyme, tone) product -------------------------
    @torch.no_grad()
    def beam_generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                      tokenized_ocr, start_symbol, end_symbol, max_len=100, num_beam=2):
        """Beam search with the key/value cache.  The reference's core class ignores `isgreedy` / `num_beam`
        (`generate` above stays greedy, like core/model/PhonemeLaTr.py:146-167); the prototype it was refactored
        from scores a triple by log p(onset) + log p(rhyme) + log p(tone) and ranks the full V_o x V_r x V_t product
        with Python loops (PhonoLaTr/ModelLaTr.py:294-377).  This is that scoring rule as a standard batched beam
        search: the joint top-k is taken exactly from the per-head top-k lists (`phoneme_beam_search`), all beams of
        all samples go through ONE cached decoder step per position, finished beams are frozen."""
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        return self._beam_from_memory(enc, attention_mask, start_symbol, end_symbol, max_len, num_beam)

    def _beam_from_memory(self, enc, attention_mask, start_symbol, end_symbol, max_len, num_beam):
        bz, K = enc.shape[0], int(num_beam)
        mem = enc.repeat_interleave(K, dim=0)
        mask = attention_mask.repeat_interleave(K, dim=0)
        cache = self.decoder.new_cache(mem, max_len + 1, self.compute_dtype)
        pe = self.positional_encoding.pos_embedding

        def step(tok, t, src):
            if src is not None:
                cache.reorder(src)
            cache.len = t
            emb = self.tgt_tok_emb(tok, pe[:, t:t + 1], out_dtype=torch.float32)
            out = self.decoder.step(emb, cache, mask, self.compute_dtype)
            # like the greedy loop (reference :195-205) the heads read the decoder output directly
            return tuple(torch.log_softmax(x[:, -1].float(), dim=-1) for x in self._heads(out.to(self.compute_dtype)))

        return phoneme_beam_search(step, bz, K, start_symbol, end_symbol, max_len, enc.device)

    def _decode_step_ids(self, tok, t, cache, attention_mask):
        """one cached greedy step: last emitted triples (B,1,3) at target position t -> next triples (B,3)"""
        emb = self.tgt_tok_emb(tok, self.positional_encoding.pos_embedding[:, t:t + 1], out_dtype=torch.float32)
        cache.len = t
        out = self.decoder.step(emb, cache, attention_mask, self.compute_dtype)
        on, rh, to = self._heads(out.to(self.compute_dtype))      # like the reference, no shared_lm_head here
        return torch.stack([on[:, -1].float().argmax(-1), rh[:, -1].float().argmax(-1), to[:, -1].float().argmax(-1)], dim=-1)

    @torch.no_grad()
    def _greedy_graphed(self, enc, attention_mask, start_symbol, end_symbol, max_len, check_every=4):
        """Cached greedy decoding with one CUDA graph per target position.  The workspace (token buffer, id
        buffer, K/V caches, mask) is address-stable and kept on the model object per (batch, memory length,
        max_len, dtype, parameter storage); graphs are captured the first time a position is reached and replayed
        afterwards.  Termination is checked every `check_every` positions and the result is cut where the
        reference loop would have stopped, so the returned ids are the reference's."""
        bz, S, _ = enc.shape
        dev = enc.device
        key = (bz, S, max_len, self.compute_dtype, str(dev), next(self.decoder.parameters()).data_ptr())
        ws = self.__dict__.get("_decode_ws")
        if ws is None or ws["key"] != key:
            ws = {"key": key, "graphs": {}, "pool": torch.cuda.graph_pool_handle(),
                  "tok": torch.zeros((bz, 1, 3), dtype=torch.long, device=dev),
                  "ys": torch.zeros((bz, max_len + 1, 3), dtype=torch.long, device=dev),
                  "mask": torch.zeros((bz, S), dtype=torch.float32, device=dev),
                  "cache": self.decoder.new_cache(enc, max_len + 1, self.compute_dtype)}
            self.__dict__["_decode_ws"] = ws
        else:
            ws["cache"].set_memory(enc.to(self.compute_dtype))
        ws["mask"].copy_(attention_mask.to(torch.float32))
        start = torch.tensor([start_symbol, 0, 0], dtype=torch.long, device=dev)
        ws["tok"].copy_(start.expand(bz, 1, 3))
        ws["ys"].zero_()
        ws["ys"][:, 0] = start

        def body(t):
            nxt = self._decode_step_ids(ws["tok"], t, ws["cache"], ws["mask"])
            ws["ys"][:, t + 1].copy_(nxt)
            ws["tok"].copy_(nxt.view(bz, 1, 3))

        n_done = max_len
        for t in range(max_len):
            g = ws["graphs"].get(t)
            if g is None:
                # eager once (lazy initialisations, kernel attributes) on a side stream, rewind, then capture
                tok0 = ws["tok"].clone()
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    body(t)
                torch.cuda.current_stream(dev).wait_stream(side)
                ws["tok"].copy_(tok0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=ws["pool"]):
                    body(t)
                ws["graphs"][t] = g
            g.replay()
            if (t + 1) % check_every == 0 or t + 1 == max_len:
                ended = ws["ys"][:, :t + 2, 0] == end_symbol
                if bool(ended.any(dim=1).all()):
                    # the reference stops after the first position at which every row has emitted <eos>
                    first = torch.where(ended, torch.arange(t + 2, device=dev)[None], t + 2).min(dim=1).values
                    n_done = int(first.max())
                    break
        return ws["ys"][:, :n_done + 1].clone()


def phoneme_beam_search(step, bz, num_beam, start_symbol, end_symbol, max_len, device):
    """Batched beam search over phoneme triples.

    step(tok (bz*K, 1, 3) int64, t, src (bz*K,) int64 or None) -> three log-prob tensors (bz*K, V_o / V_r / V_t) for
    target position t + 1; `src[n]` names the row (of the previous call) whose state row n continues.
    Score of a hypothesis = sum over positions of log p(onset) + log p(rhyme) + log p(tone).  Because that sum is
    separable, the K best triples of one beam lie in the product of the K best entries of each head, so the joint
    top-K over all K * V_o * V_r * V_t continuations is found exactly among K * K^3 candidates.  A beam that has
    emitted `end_symbol` (onset column) is finished: it keeps its score and is extended by (end_symbol, 0, 0) only.
    Returns (bz, L, 3): the best hypothesis per sample, L <= max_len + 1, positions after <eos> hold (end_symbol, 0, 0).
    Ties are broken by torch.topk's order (lower flat index first)."""
    K = num_beam
    NEG = float("-inf")
    seqs = torch.zeros((bz, K, 1, 3), dtype=torch.long, device=device)
    seqs[:, :, 0, 0] = start_symbol
    scores = torch.full((bz, K), NEG, device=device)
    scores[:, 0] = 0.0                                   # all beams start identical: only beam 0 may expand
    finished = torch.zeros((bz, K), dtype=torch.bool, device=device)
    src = None
    base = (torch.arange(bz, device=device) * K)[:, None]
    for t in range(max_len):
        on, rh, to = step(seqs[:, :, -1].reshape(bz * K, 1, 3), t, src)
        k_o, k_r, k_t = min(K, on.shape[-1]), min(K, rh.shape[-1]), min(K, to.shape[-1])
        ov, oi = on.view(bz, K, -1).topk(k_o, dim=-1)
        rv, ri = rh.view(bz, K, -1).topk(k_r, dim=-1)
        tv, ti = to.view(bz, K, -1).topk(k_t, dim=-1)
        joint = ov[:, :, :, None, None] + rv[:, :, None, :, None] + tv[:, :, None, None, :]     # (bz,K,ko,kr,kt)
        # finished beams: one continuation, (end, 0, 0), at no cost
        only_first = torch.full_like(joint, NEG)
        only_first[:, :, 0, 0, 0] = 0.0
        joint = torch.where(finished[:, :, None, None, None], only_first, joint)
        cand = (scores[:, :, None, None, None] + joint).reshape(bz, -1)
        scores, flat = cand.topk(K, dim=-1)
        per_beam = k_o * k_r * k_t
        beam = flat // per_beam
        rem = flat % per_beam
        a, b_, c = rem // (k_r * k_t), (rem // k_t) % k_r, rem % k_t
        was_finished = finished.gather(1, beam)
        new_o = oi[torch.arange(bz, device=device)[:, None], beam, a]
        new_r = ri[torch.arange(bz, device=device)[:, None], beam, b_]
        new_t = ti[torch.arange(bz, device=device)[:, None], beam, c]
        new_o = torch.where(was_finished, torch.full_like(new_o, end_symbol), new_o)
        new_r = torch.where(was_finished, torch.zeros_like(new_r), new_r)
        new_t = torch.where(was_finished, torch.zeros_like(new_t), new_t)
        nxt = torch.stack([new_o, new_r, new_t], dim=-1)                                          # (bz,K,3)
        seqs = torch.cat([seqs.gather(1, beam[:, :, None, None].expand(-1, -1, seqs.shape[2], 3)), nxt[:, :, None]], dim=2)
        finished = was_finished | (new_o == end_symbol)
        src = (base + beam).reshape(-1)
        if bool(finished.all()):
            break
    best = scores.argmax(dim=-1)
    return seqs[torch.arange(bz, device=device), best]


class PhonemePreSTU(nn.Module, _VisionMixin):
    """reference: core/model/PhonemePreSTU.py:16-199 (intended behaviour).
    No layout branch, ViT NOT frozen."""

    def __init__(self, config, onset_vocab_size, rhyme_vocab_size, tone_vocab_size):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        d = config.d_model
        self.rhyme_tone_embed_dim = d // 3
        self.onset_embed_dim = int(d - self.rhyme_tone_embed_dim * 2)
        self.tgt_tok_emb = PhonemeEmbedding(onset_vocab_size, rhyme_vocab_size, tone_vocab_size,
                                            self.onset_embed_dim, self.rhyme_tone_embed_dim)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(emb_size=d, num_layers=config.num_decoder_layers, n_head=config.n_head)
        self.shared_lm_head = nn.Linear(d, d)
        self.onset_lm_head = nn.Linear(self.onset_embed_dim, onset_vocab_size)
        self.rhyme_lm_head = nn.Linear(self.rhyme_tone_embed_dim, rhyme_vocab_size)
        self.tone_lm_head = nn.Linear(self.rhyme_tone_embed_dim, tone_vocab_size)

    def _calculate_embedding(self, pixel_values, input_ids, src_attention_mask):
        vit_tokens = self._vit_tokens(pixel_values)
        img_feat = _lin(vit_tokens.to(self.compute_dtype), self.visual_projector.weight, self.visual_projector.bias)
        return ops.embed_multimodal(img_feat, None, None, input_ids, None, src_attention_mask,
                                    self.encoder.shared.weight, (), out_dtype=self.compute_dtype)

    calculate_embedding = _calculate_embedding   # the reference's forward calls this name (D5)

    decode = PhonemeLaTr.decode
    _heads = PhonemeLaTr._heads

    def forward(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask):
        inputs_embeds, attention_mask = self._calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        dec = self.decode(labels, enc, attention_mask, label_attention_mask)
        h = _lin(dec.to(self.compute_dtype), self.shared_lm_head.weight, self.shared_lm_head.bias)
        on, rh, to = self._heads(h)
        return on.float(), rh.float(), to.float()

    def forward_loss(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask, targets,
                     ignore_index):
        """model forward + the executor's 3x CrossEntropyLoss (core/executor/PhonemePreSTU_Executor.py:160-180)
        through the fused phoneme head kernel; `targets` = labels[:, 1:, :] of the executor."""
        inputs_embeds, attention_mask = self._calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        dec = self.decode(labels, enc, attention_mask, label_attention_mask)
        return _phoneme_head_loss(self, dec.to(self.compute_dtype), targets, ignore_index)

    # The call the executor makes (core/executor/PhonemePreSTU_Executor.py:41-49): the class's own `generate`
    # (PhonemePreSTU.py:103-199) still carries PhonemeLaTr's argument list and cannot run, so the
    # intended behaviour is PhonemeLaTr's greedy loop on this model's encoder; isgreedy / num_beam are ignored there.
    def generate(self, pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_length=20,
                 isgreedy=True, num_beam=2):
        return self.greedy_generate(pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_length)

    _beam_from_memory = PhonemeLaTr._beam_from_memory

    @torch.no_grad()
    def beam_generate(self, pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_len=100,
                      num_beam=2):
        """factorised beam search with the key/value cache (see PhonemeLaTr.beam_generate)"""
        inputs_embeds, attention_mask = self._calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        return self._beam_from_memory(enc, attention_mask, start_symbol, end_symbol, max_len, num_beam)

    @torch.no_grad()
    def greedy_generate(self, pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_len=100,
                        use_cache=True):
        bz, dev = input_ids.size(0), input_ids.device
        inputs_embeds, attention_mask = self._calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        ys = torch.tensor([[[start_symbol, 0, 0]]], dtype=torch.long, device=dev).repeat(bz, 1, 1)
        cache = self.decoder.new_cache(enc, max_len + 1, self.compute_dtype) if use_cache else None
        pe = self.positional_encoding.pos_embedding
        for t in range(max_len):
            if use_cache:
                emb = self.tgt_tok_emb(ys[:, -1:], pe[:, t:t + 1], out_dtype=torch.float32)
                out = self.decoder.step(emb, cache, attention_mask, self.compute_dtype)
            else:
                out = self.decode(ys, enc, attention_mask)[:, -1:]
            on, rh, to = self._heads(out.to(self.compute_dtype))       # no shared_lm_head, like PhonemeLaTr.py:195-205
            nxt = torch.stack([on[:, -1].float().argmax(-1), rh[:, -1].float().argmax(-1),
                               to[:, -1].float().argmax(-1)], dim=-1)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
            if torch.any(ys[:, :, 0] == end_symbol, dim=1).sum() == bz:
                break
        return ys


def _load_pretrained_t5(backbone: T5ForConditionalGeneration, config):
    if _random_init(config):
        return
    from transformers import T5ForConditionalGeneration as HF
    backbone.load_state_dict(HF.from_pretrained(config._name_or_path).state_dict(), strict=True)


class LaTr(nn.Module, _VisionMixin):
    """reference: core/model/LaTr.py:42-111 — T5 encoder + T5 decoder + lm_head over the T5 vocabulary."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.backbone = T5ForConditionalGeneration(config)
        _load_pretrained_t5(self.backbone, config)
        self.spatial_feat_extractor = SpatialModule(config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        for _, child in self.vit.named_children():
            for param in child.parameters():
                param.requires_grad = False

    # reference :85-97
    def calculate_embedding(self, pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask,
                            tokenized_ocr):
        vit_tokens = self._vit_tokens(pixel_values)
        img_feat = _lin(vit_tokens.to(self.compute_dtype), self.visual_projector.weight, self.visual_projector.bias)
        return ops.embed_multimodal(img_feat, coordinates, tokenized_ocr, input_ids, ocr_attention_mask,
                                    src_attention_mask, self.backbone.shared.weight,
                                    self.spatial_feat_extractor.tables(), out_dtype=self.compute_dtype)

    def _decoder_hidden(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                        ocr_attention_mask, tokenized_ocr):
        inputs_embeds, attention_mask = self.calculate_embedding(
            pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask, tokenized_ocr)
        enc = self.backbone.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        # reference :76-80: decoder gets NO encoder_attention_mask; label mask is 1 = valid (int64)
        tgt = torch.nn.functional.embedding(labels, self.backbone.shared.weight)
        return self.backbone.decoder(tgt, label_attention_mask, compute_dtype=self.compute_dtype, memory=enc,
                                     memory_mask=None)

    # reference :58-83
    def forward(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                ocr_attention_mask, tokenized_ocr):
        dec = self._decoder_hidden(pixel_values, coordinates, input_ids, labels, src_attention_mask,
                                   label_attention_mask, ocr_attention_mask, tokenized_ocr)
        # lm_head is called directly: no d_model**-0.5 rescale (reference :83)
        return _lin(dec.to(self.compute_dtype), self.backbone.lm_head.weight).float()

    def forward_loss(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                     ocr_attention_mask, tokenized_ocr, targets, ignore_index):
        """model forward + CrossEntropyLoss(ignore_index=pad) of core/executor/LaTr_Executor.py:160-163."""
        dec = self._decoder_hidden(pixel_values, coordinates, input_ids, labels, src_attention_mask,
                                   label_attention_mask, ocr_attention_mask, tokenized_ocr)
        h = dec.to(self.compute_dtype).reshape(-1, dec.shape[-1])
        w = self.backbone.lm_head.weight
        w_lp = SHADOWS.get([w], h.dtype) if h.dtype != w.dtype else None
        return ops.vocab_head_ce(h, w, targets, ignore_index, w_lp=w_lp)

    # reference :99-111 — HF greedy generation from inputs_embeds (no encoder attention mask is passed there)
    @torch.no_grad()
    def generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask, tokenized_ocr,
                 max_length=20):
        inputs_embeds, _ = self.calculate_embedding(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                                    src_attention_mask, tokenized_ocr)
        enc = self.backbone.encoder(inputs_embeds, None, compute_dtype=self.compute_dtype)
        return _t5_greedy(self.backbone, self.config, enc, max_length, self.compute_dtype)


def _t5_greedy(backbone, cfg, enc, max_length, compute_dtype):
    """HF `generate(inputs_embeds=..., max_length=...)` with the default greedy settings: start from
    decoder_start_token_id, at most max_length ids per row, finished rows are padded with pad_token_id, stop when
    every row has emitted eos_token_id.  The decoder gets no encoder mask (the reference passes none)."""
    start = cfg.decoder_start_token_id if cfg.decoder_start_token_id is not None else cfg.pad_token_id
    B = enc.shape[0]
    ys = torch.full((B, 1), start, dtype=torch.long, device=enc.device)
    done = torch.zeros(B, dtype=torch.bool, device=ys.device)
    for _ in range(max_length - 1):
        tgt = torch.nn.functional.embedding(ys, backbone.shared.weight)
        dec = backbone.decoder(tgt, None, compute_dtype=compute_dtype, memory=enc, memory_mask=None)
        logits = _lin(dec[:, -1:].to(compute_dtype), backbone.lm_head.weight).float()
        nxt = logits[:, -1].argmax(-1)
        nxt = torch.where(done, torch.full_like(nxt, cfg.pad_token_id), nxt)
        ys = torch.cat([ys, nxt[:, None]], dim=1)
        done = done | (nxt == cfg.eos_token_id)
        if bool(done.all()):
            break
    return ys


# reference: core/model/PhonemeSaL.py:15-25
class CustomizedSaL_config:
    def build(self, config, new_token_embedding_size):
        model_config = _auto_config(config.backbone_name)
        model_config.update({"ocr_hidden": config.ocr_hidden, "obj_hidden": config.obj_hidden,
                             "new_token_embedding_size": new_token_embedding_size,
                             "num_decoder_layers": config.num_decoder_layers, "n_head": config.n_head})
        return model_config


class PhonemeSaL(nn.Module):
    """reference: core/model/PhonemeSaL.py:28-207.  question ‖ OCR(det+rec features, box, token) ‖ objects ->
    T5 encoder with EXTERNAL 1-D + SCP position bias (the attention mask is then NOT applied: HF only adds it
    when it computes the bias itself — SURVEY D14) -> 4-layer target decoder over the flat 253-phoneme
    vocabulary -> lm_head, CrossEntropyLoss(ignore_index=0) inside the model."""

    def __init__(self, config, vocab_size, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        self.config = config
        self.vocab_size = vocab_size
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)
        new_size = getattr(config, "new_token_embedding_size", None)
        if new_size is not None and new_size != self.encoder.shared.num_embeddings:
            self._resize_token_embeddings(new_size)
        self.rel2Dbias = RelativePositionBiasAggregated(
            Relative1D=RelativePositionBias1D(num_heads=config.num_heads),
            SCP=SCPRelativePositionBias(num_heads=config.num_heads))
        d = config.d_model
        self.obj_dropout = nn.Dropout(obj_dropout)          # declared by the reference, never applied (:44,:49)
        self.obj_feature_projector = nn.Linear(config.obj_hidden, d)
        self.obj_bbox_projector = nn.Linear(4, d)
        self.obj_feature_layer_norm = T5LayerNorm(d)
        self.ocr_dropout = nn.Dropout(ocr_dropout)
        self.ocr_feature_projector = nn.Linear(config.ocr_hidden, d)
        self.ocr_bbox_projector = nn.Linear(4, d)
        self.ocr_feature_layer_norm = T5LayerNorm(d)
        self.tgt_tok_emb = nn.Embedding(num_embeddings=vocab_size, embedding_dim=d)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(emb_size=d, num_layers=config.num_decoder_layers, n_head=config.n_head)
        self.lm_head = nn.Linear(d, vocab_size)
        self.loss_fn = nn.CrossEntropyLoss(ignore_index=0)

    def set_compute_dtype(self, dtype):
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        return self

    def _resize_token_embeddings(self, n):
        old = self.encoder.shared
        new = nn.Embedding(n, old.embedding_dim)
        nn.init.normal_(new.weight, mean=0.0, std=self.config.initializer_factor)
        k = min(n, old.num_embeddings)
        new.weight.data[:k] = old.weight.data[:k]
        self.encoder.shared = new
        self.encoder.encoder.embed_tokens = new

    # reference :194-202 — the same T5LayerNorm module normalises both the feature and the box projection
    def _modality_embedding(self, tokens, coords, feats, feat_proj, box_proj, norm):
        cd = self.compute_dtype
        a = norm(_lin(feats.to(cd), feat_proj.weight, feat_proj.bias).float(), out_dtype=torch.float32)
        b = norm(_lin(coords.to(cd), box_proj.weight, box_proj.bias).float(), out_dtype=torch.float32)
        return a + b + torch.nn.functional.embedding(tokens, self.encoder.shared.weight)

    def _calculate_obj_embedding(self, tokenized_obj, obj_coordinates, obj_features):
        return self._modality_embedding(tokenized_obj, obj_coordinates, obj_features, self.obj_feature_projector,
                                        self.obj_bbox_projector, self.obj_feature_layer_norm)

    def _calculate_ocr_embedding(self, tokenized_ocr, ocr_coordinates, ocr_features):
        return self._modality_embedding(tokenized_ocr, ocr_coordinates, ocr_features, self.ocr_feature_projector,
                                        self.ocr_bbox_projector, self.ocr_feature_layer_norm)

    def _encode(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates, ocr_features,
                tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques):
        obj = self._calculate_obj_embedding(tokenized_obj, obj_coordinates, obj_features)
        ocr = self._calculate_ocr_embedding(tokenized_ocr, ocr_coordinates, ocr_features)
        ques = torch.nn.functional.embedding(input_ids, self.encoder.shared.weight)
        feat = torch.cat([ques, ocr, obj], dim=1)
        mask = torch.cat([src_attention_mask, ocr_attention_mask, obj_attention_mask], dim=1)
        rel, scp = self.rel2Dbias(feat.shape[1], ocr_coordinates, int(max_ques), int(max_ocr))
        enc = self.encoder.encoder(feat, None, compute_dtype=self.compute_dtype, external_rel_bias=rel, scp=scp)
        return enc, mask

    # reference :122-131
    def decode(self, labels, encoder_outputs, encoder_attention_mask, label_attention_mask=None):
        emb = self.positional_encoding(torch.nn.functional.embedding(labels, self.tgt_tok_emb.weight))
        return self.decoder(emb, encoder_outputs, memory_key_padding_mask=encoder_attention_mask,
                            tgt_key_padding_mask=label_attention_mask, compute_dtype=self.compute_dtype, causal=True)

    # reference :73-120
    def forward(self, input_ids, src_attention_mask, label_ids, shifted_right_label_ids, label_attention_mask,
                tokenized_ocr, ocr_attention_mask, ocr_coordinates, ocr_features, tokenized_obj, obj_attention_mask,
                obj_coordinates, obj_features, max_ocr, max_ques):
        enc, mask = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features,
                                 max_ocr, max_ques)
        dec = self.decode(label_ids, enc, mask, label_attention_mask)
        logits = _lin(dec.to(self.compute_dtype), self.lm_head.weight, self.lm_head.bias).float()
        loss = self.loss_fn(logits.reshape((-1, self.vocab_size)), shifted_right_label_ids.reshape(-1))
        return logits, loss

    # reference :134-192
    @torch.no_grad()
    def generate(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques,
                 start_symbol, end_symbol, max_len=100):
        bz = input_ids.size(0)
        enc, mask = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features,
                                 max_ocr, max_ques)
        ys = torch.full((bz, 1), start_symbol, dtype=torch.long, device=input_ids.device)
        done = torch.zeros_like(ys)
        for _ in range(max_len):
            out = self.decode(ys, enc, mask)
            logits = _lin(out[:, -1:].to(self.compute_dtype), self.lm_head.weight, self.lm_head.bias).float()
            nxt = logits[:, -1].argmax(-1)
            done = torch.where((nxt == end_symbol)[:, None], torch.ones_like(done), done)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
            if bool(done.all()):
                break
        return ys


# ----------------------------------------------------------------------------------
# the rest of the family: PreSTU / SaL (T5 encoder-decoder over the T5 vocabulary) and the Customized* classes
# (T5 encoder + 4-layer target decoder over a flat char / byte / BPE vocabulary).  They are compositions of the
# same kernels: K1 (fused multimodal embedding), K2 (T5 attention, with the in-kernel SCP bias for SaL), K3
# (target decoder attention with additive float masks), the fused norm chains, and cuBLAS linears.
# ----------------------------------------------------------------------------------
def _resize_shared(holder, stack_owner, n, config):
    """HF resize_token_embeddings on a model whose `shared` table is tied to the stacks' embed_tokens
    (core/model/SaL.py:30, PhonemeSaL.py:41): first min(n, old) rows kept, new rows drawn like HF's init."""
    old = holder.shared
    if n is None or n == old.num_embeddings:
        return
    new = nn.Embedding(n, old.embedding_dim)
    nn.init.normal_(new.weight, mean=0.0, std=config.initializer_factor)
    k = min(n, old.num_embeddings)
    new.weight.data[:k] = old.weight.data[:k]
    holder.shared = new
    for st in stack_owner:
        st.embed_tokens = new


class PreSTU(nn.Module, _VisionMixin):
    """reference: core/model/PreSTU.py:13-66 — ViT tokens ‖ (question + OCR text) -> T5 encoder-decoder -> lm_head.
    No layout branch; the ViT is NOT frozen (no freeze loop in the reference)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.backbone = T5ForConditionalGeneration(config)
        _load_pretrained_t5(self.backbone, config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)

    # reference :48-56
    def calculate_embedding(self, pixel_values, input_ids, src_attention_mask):
        vit_tokens = self._vit_tokens(pixel_values)
        img_feat = _lin(vit_tokens.to(self.compute_dtype), self.visual_projector.weight, self.visual_projector.bias)
        return ops.embed_multimodal(img_feat, None, None, input_ids, None, src_attention_mask,
                                    self.backbone.shared.weight, (), out_dtype=self.compute_dtype)

    def _decoder_hidden(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask):
        inputs_embeds, attention_mask = self.calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.backbone.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype)
        tgt = torch.nn.functional.embedding(labels, self.backbone.shared.weight)
        return self.backbone.decoder(tgt, label_attention_mask, compute_dtype=self.compute_dtype, memory=enc,
                                     memory_mask=None)

    # reference :24-46
    def forward(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask):
        dec = self._decoder_hidden(pixel_values, input_ids, labels, src_attention_mask, label_attention_mask)
        return _lin(dec.to(self.compute_dtype), self.backbone.lm_head.weight).float()

    def forward_loss(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask, targets,
                     ignore_index):
        """model forward + CrossEntropyLoss(ignore_index=pad) of core/executor/PreSTU_Executor.py without the
        (N, V) logits (chunked lm_head + fused softmax / CE / dlogits)."""
        dec = self._decoder_hidden(pixel_values, input_ids, labels, src_attention_mask, label_attention_mask)
        h = dec.to(self.compute_dtype).reshape(-1, dec.shape[-1])
        w = self.backbone.lm_head.weight
        w_lp = SHADOWS.get([w], h.dtype) if h.dtype != w.dtype else None
        return ops.vocab_head_ce(h, w, targets, ignore_index, w_lp=w_lp)

    # reference :58-66
    @torch.no_grad()
    def generate(self, pixel_values, input_ids, src_attention_mask, max_length=20):
        inputs_embeds, _ = self.calculate_embedding(pixel_values, input_ids, src_attention_mask)
        enc = self.backbone.encoder(inputs_embeds, None, compute_dtype=self.compute_dtype)
        return _t5_greedy(self.backbone, self.config, enc, max_length, self.compute_dtype)


class _FlatTargetMixin:
    """Target side shared by CustomizedLaTr / CustomizedPreSTU / CustomizedSaL: TokenEmbedding (x sqrt(d)) ->
    sinusoidal PE + dropout -> BaseDecoder -> lm_head (with bias) over a flat vocabulary; greedy decoding (with a
    key/value cache) and the reference's beam routine."""

    def _init_flat_target(self, config, tgt_vocab_size):
        d = config.d_model
        self.tgt_tok_emb = TokenEmbedding(tgt_vocab_size, d)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(emb_size=d, num_layers=config.num_decoder_layers, n_head=config.n_head)
        self.lm_head = nn.Linear(d, tgt_vocab_size)

    # CustomizedLaTr.py:99-110, CustomizedPreSTU.py:60-70, CustomizedSaL.py:108-118
    def decode(self, labels, encoder_outputs, encoder_attention_mask, label_attention_mask=None):
        emb = self.positional_encoding(self.tgt_tok_emb(labels))
        return self.decoder(emb, encoder_outputs, tgt_mask=None, memory_key_padding_mask=encoder_attention_mask,
                            tgt_key_padding_mask=label_attention_mask, compute_dtype=self.compute_dtype, causal=True)

    def _logits(self, dec):
        return _lin(dec.to(self.compute_dtype), self.lm_head.weight, self.lm_head.bias).float()

    @staticmethod
    def _memory_mask_for_step(mask):
        if mask is not None and mask.dtype == torch.bool:        # bool masks mask, float masks are additive (D14)
            return torch.where(mask, float("-inf"), 0.0).to(torch.float32)
        return mask

    # CustomizedLaTr.py:146-183 — `use_cache=False` is the reference's O(T^2) loop, token for token
    @torch.no_grad()
    def _greedy(self, enc, attention_mask, start_symbol, end_symbol, max_len, use_cache=True):
        bz, dev = enc.shape[0], enc.device
        ys = torch.full((bz, 1), start_symbol, dtype=torch.long, device=dev)
        cache = self.decoder.new_cache(enc, max_len + 1, self.compute_dtype) if use_cache else None
        step_mask = self._memory_mask_for_step(attention_mask)
        pe = self.positional_encoding.pos_embedding
        for t in range(max_len):
            if use_cache:
                emb = self.tgt_tok_emb(ys[:, -1:]) + pe[:, t:t + 1]
                out = self.decoder.step(emb, cache, step_mask, self.compute_dtype)
            else:
                out = self.decode(ys, enc, attention_mask)[:, -1:]
            nxt = self._logits(out)[:, -1].argmax(-1).view(bz, 1)
            ys = torch.cat([ys, nxt], dim=1)
            if torch.any(ys == end_symbol, dim=1).sum() == bz:
                break
        return ys

    # CustomizedLaTr.py:185-249, CustomizedSaL.py:235-320
    @torch.no_grad()
    def _beam(self, enc, attention_mask, start_symbol, end_symbol, max_len, num_beam):
        ys = torch.full((enc.shape[0], 1), start_symbol, dtype=torch.long, device=enc.device)
        prob = self._logits(self.decode(ys, enc, attention_mask)[:, -1:])[:, -1]
        return reference_beam_select(prob, ys, end_symbol, max_len, num_beam)


def reference_beam_select(prob, ys, end_symbol, max_len, num_beam):
    """The reference's `beam_generate` after its first decoder call, restated on that call's scores `prob` (B, V).
    The routine re-decodes the ONE-token prefix `ys` for every beam and every step (`self.decode(ys, ...)`, never
    `beams[b]`), so every later score equals `prob`: each beam is its first-step candidate followed by the arg-max
    token repeated.  Kept as the reference computes it — `log` of raw scores (NaN for negative ones, and torch's
    argmax picks a NaN entry), one `eos_mask` tensor shared by all beams (`[mask] * num_beam`), per-beam early
    exit — because the returned ids are the contract; the invariant decoder call is simply done once."""
    bz = ys.shape[0]
    values, indices = torch.topk(prob, num_beam, dim=-1)
    beams = [torch.cat([ys, indices[:, i:i + 1]], dim=1) for i in range(num_beam)]
    beam_probs = [torch.log(values[:, i:i + 1]) for i in range(num_beam)]
    vals, inds = values[:, :1], indices[:, :1]           # topk(prob, 1) of the re-decoded one-token prefix
    log_vals = torch.log(vals)
    done = [False] * num_beam
    eos_mask = torch.ones((bz, 1), dtype=torch.long, device=prob.device)      # shared by every beam
    for _ in range(max_len - 1):
        for b in range(num_beam):
            eos_mask = eos_mask * (inds != end_symbol)
            beams[b] = torch.cat([beams[b], inds], dim=1)
            if eos_mask.sum() == 0:
                done[b] = True
                continue
            beam_probs[b] = beam_probs[b] + log_vals * eos_mask
        if all(done):
            break
    # like the reference, the selection itself happens on the host (CustomizedLaTr.py:241-247): torch's CPU argmax
    # semantics for NaN scores are then the reference's by construction
    beam_probs = torch.cat(beam_probs, dim=-1).cpu()
    beams = torch.stack(beams, dim=1).cpu()
    beam_idx = torch.argmax(beam_probs, dim=-1)
    return beams[torch.arange(bz), beam_idx.flatten(), :]


class CustomizedLaTr(nn.Module, _VisionMixin, _FlatTargetMixin):
    """reference: core/model/CustomizedLaTr.py:45-271"""

    def __init__(self, config, tgt_vocab_size=300):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)
        self.spatial_feat_extractor = SpatialModule(config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        for _, child in self.vit.named_children():                 # reference :56-59
            for param in child.parameters():
                param.requires_grad = False
        self._init_flat_target(config, tgt_vocab_size)

    _calculate_embedding = PhonemeLaTr._calculate_embedding       # reference :251-264, same body
    _encode = PhonemeLaTr._encode

    # reference :73-97
    def forward(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                ocr_attention_mask, tokenized_ocr):
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        return self._logits(self.decode(labels, enc, attention_mask, label_attention_mask))

    # reference :112-144
    def generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask, tokenized_ocr,
                 start_symbol, end_symbol, max_length=20, isgreedy=True, num_beam=2):
        if isgreedy:
            return self.greedy_generate(pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                                        tokenized_ocr, start_symbol, end_symbol, max_length)
        return self.beam_generate(pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                                  tokenized_ocr, start_symbol, end_symbol, max_length, num_beam)

    @torch.no_grad()
    def greedy_generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                        tokenized_ocr, start_symbol, end_symbol, max_len=100, use_cache=True):
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        return self._greedy(enc, attention_mask, start_symbol, end_symbol, max_len, use_cache)

    @torch.no_grad()
    def beam_generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                      tokenized_ocr, start_symbol, end_symbol, max_len=100, num_beam=2):
        enc, attention_mask = self._encode(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                           src_attention_mask, tokenized_ocr)
        return self._beam(enc, attention_mask, start_symbol, end_symbol, max_len, num_beam)


class CustomizedPreSTU(nn.Module, _VisionMixin, _FlatTargetMixin):
    """reference: core/model/CustomizedPreSTU.py:16-143.  No layout branch, ViT NOT frozen."""

    def __init__(self, config, tgt_vocab_size):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)
        self.vit = _build_vit(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        self._init_flat_target(config, tgt_vocab_size)

    _calculate_embedding = PhonemePreSTU._calculate_embedding     # reference :127-136, same body

    def _encode(self, pixel_values, input_ids, src_attention_mask):
        inputs_embeds, attention_mask = self._calculate_embedding(pixel_values, input_ids, src_attention_mask)
        return self.encoder.encoder(inputs_embeds, attention_mask, compute_dtype=self.compute_dtype), attention_mask

    # reference :37-58
    def forward(self, pixel_values, input_ids, labels, src_attention_mask, label_attention_mask):
        enc, attention_mask = self._encode(pixel_values, input_ids, src_attention_mask)
        return self._logits(self.decode(labels, enc, attention_mask, label_attention_mask))

    # reference :72-87 (isgreedy / num_beam are accepted and ignored there too)
    def generate(self, pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_length=20,
                 isgreedy=True, num_beam=2):
        return self.greedy_generate(pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_length)

    @torch.no_grad()
    def greedy_generate(self, pixel_values, input_ids, src_attention_mask, start_symbol, end_symbol, max_len=100,
                        use_cache=True):
        enc, attention_mask = self._encode(pixel_values, input_ids, src_attention_mask)
        return self._greedy(enc, attention_mask, start_symbol, end_symbol, max_len, use_cache)


class _SaLInputs:
    """question ‖ OCR ‖ object embedding of the SaL family (SaL.py:93-101, CustomizedSaL.py:323-331): the same
    T5LayerNorm normalises the region-feature projection and the box projection, the token embedding is added."""

    def _init_sal_inputs(self, config, num_heads, obj_dropout, ocr_dropout):
        d = config.d_model
        self.rel2Dbias = RelativePositionBiasAggregated(Relative1D=RelativePositionBias1D(num_heads=num_heads),
                                                        SCP=SCPRelativePositionBias(num_heads=num_heads))
        self.obj_dropout = nn.Dropout(obj_dropout)          # declared by the reference, never applied
        self.obj_feature_projector = nn.Linear(config.obj_hidden, d)
        self.obj_bbox_projector = nn.Linear(4, d)
        self.obj_feature_layer_norm = T5LayerNorm(d)
        self.ocr_dropout = nn.Dropout(ocr_dropout)
        self.ocr_feature_projector = nn.Linear(config.ocr_hidden, d)
        self.ocr_bbox_projector = nn.Linear(4, d)
        self.ocr_feature_layer_norm = T5LayerNorm(d)

    def _modality_embedding(self, shared, tokens, coords, feats, feat_proj, box_proj, norm):
        cd = self.compute_dtype
        a = norm(_lin(feats.to(cd), feat_proj.weight, feat_proj.bias).float(), out_dtype=torch.float32)
        b = norm(_lin(coords.to(cd), box_proj.weight, box_proj.bias).float(), out_dtype=torch.float32)
        return a + b + torch.nn.functional.embedding(tokens, shared.weight)

    def _sal_features(self, shared, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                      ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr,
                      max_ques):
        obj = self._modality_embedding(shared, tokenized_obj, obj_coordinates, obj_features,
                                       self.obj_feature_projector, self.obj_bbox_projector,
                                       self.obj_feature_layer_norm)
        ocr = self._modality_embedding(shared, tokenized_ocr, ocr_coordinates, ocr_features,
                                       self.ocr_feature_projector, self.ocr_bbox_projector,
                                       self.ocr_feature_layer_norm)
        ques = torch.nn.functional.embedding(input_ids, shared.weight)
        feat = torch.cat([ques, ocr, obj], dim=1)
        mask = torch.cat([src_attention_mask, ocr_attention_mask, obj_attention_mask], dim=1)
        rel, scp = self.rel2Dbias(feat.shape[1], ocr_coordinates, int(max_ques), int(max_ocr))
        return feat, mask, rel, scp


class SaL(nn.Module, _SaLInputs):
    """reference: core/model/SaL.py:24-140 — T5 encoder-decoder (`T52dForConditionalGeneration`) whose encoder takes
    the EXTERNAL 1-D + SCP position bias (so the attention mask is not applied there — SURVEY D14), T5 decoder
    without an encoder mask, lm_head over the (resized) T5 vocabulary."""

    def __init__(self, config, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.backbone = T5ForConditionalGeneration(config)
        _load_pretrained_t5(self.backbone, config)
        self._resize_token_embeddings(getattr(config, "new_token_embedding_size", None))
        self._init_sal_inputs(config, config.num_heads, obj_dropout, ocr_dropout)

    def set_compute_dtype(self, dtype):
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        return self

    def _resize_token_embeddings(self, n):
        bb = self.backbone
        if n is None or n == bb.shared.num_embeddings:
            return
        tied = bb.lm_head.weight is bb.shared.weight
        _resize_shared(bb, (bb.encoder, bb.decoder), n, self.config)
        if tied:
            bb.lm_head = nn.Linear(bb.shared.embedding_dim, n, bias=False)
            bb.lm_head.weight = bb.shared.weight
        else:
            old = bb.lm_head
            bb.lm_head = nn.Linear(old.in_features, n, bias=False)
            nn.init.normal_(bb.lm_head.weight, mean=0.0, std=self.config.initializer_factor)
            k = min(n, old.out_features)
            bb.lm_head.weight.data[:k] = old.weight.data[:k]
        self.config.vocab_size = n

    # reference :93-101
    def calculate_obj_embedding(self, tokenized_obj, obj_coordinates, obj_features):
        return self._modality_embedding(self.backbone.shared, tokenized_obj, obj_coordinates, obj_features,
                                        self.obj_feature_projector, self.obj_bbox_projector,
                                        self.obj_feature_layer_norm)

    def calculate_ocr_embedding(self, tokenized_ocr, ocr_coordinates, ocr_features):
        return self._modality_embedding(self.backbone.shared, tokenized_ocr, ocr_coordinates, ocr_features,
                                        self.ocr_feature_projector, self.ocr_bbox_projector,
                                        self.ocr_feature_layer_norm)

    def _encode(self, *inputs):
        feat, mask, rel, scp = self._sal_features(self.backbone.shared, *inputs)
        return self.backbone.encoder(feat, None, compute_dtype=self.compute_dtype, external_rel_bias=rel, scp=scp)

    def _decoder_hidden(self, input_ids, src_attention_mask, label_ids, label_attention_mask, tokenized_ocr,
                        ocr_attention_mask, ocr_coordinates, ocr_features, tokenized_obj, obj_attention_mask,
                        obj_coordinates, obj_features, max_ocr, max_ques):
        enc = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                           ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr,
                           max_ques)
        tgt = torch.nn.functional.embedding(label_ids, self.backbone.shared.weight)
        return self.backbone.decoder(tgt, label_attention_mask, compute_dtype=self.compute_dtype, memory=enc,
                                     memory_mask=None)

    # reference :44-91
    def forward(self, input_ids, src_attention_mask, label_ids, label_attention_mask, tokenized_ocr,
                ocr_attention_mask, ocr_coordinates, ocr_features, tokenized_obj, obj_attention_mask,
                obj_coordinates, obj_features, max_ocr, max_ques):
        dec = self._decoder_hidden(input_ids, src_attention_mask, label_ids, label_attention_mask, tokenized_ocr,
                                   ocr_attention_mask, ocr_coordinates, ocr_features, tokenized_obj,
                                   obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques)
        return _lin(dec.to(self.compute_dtype), self.backbone.lm_head.weight).float()

    # reference :104-140
    @torch.no_grad()
    def generate(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques,
                 max_length=20):
        enc = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                           ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr,
                           max_ques)
        return _t5_greedy(self.backbone, self.config, enc, max_length, self.compute_dtype)


class CustomizedSaL(nn.Module, _SaLInputs, _FlatTargetMixin):
    """reference: core/model/CustomizedSaL.py:29-335 — the PhonemeSaL encoder side with the flat-vocabulary
    target decoder; forward returns logits only (the executor owns the loss)."""

    def __init__(self, config, tgt_vocab_size, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        self.config = config
        self.compute_dtype = torch.float32
        self.encoder = T5EncoderModel(config)
        _load_pretrained_t5_encoder(self.encoder, config)
        _resize_shared(self.encoder, (self.encoder.encoder,), getattr(config, "new_token_embedding_size", None),
                       config)
        self._init_sal_inputs(config, config.num_heads, obj_dropout, ocr_dropout)
        self._init_flat_target(config, tgt_vocab_size)

    set_compute_dtype = SaL.set_compute_dtype

    def _calculate_obj_embedding(self, tokenized_obj, obj_coordinates, obj_features):
        return self._modality_embedding(self.encoder.shared, tokenized_obj, obj_coordinates, obj_features,
                                        self.obj_feature_projector, self.obj_bbox_projector,
                                        self.obj_feature_layer_norm)

    def _calculate_ocr_embedding(self, tokenized_ocr, ocr_coordinates, ocr_features):
        return self._modality_embedding(self.encoder.shared, tokenized_ocr, ocr_coordinates, ocr_features,
                                        self.ocr_feature_projector, self.ocr_bbox_projector,
                                        self.ocr_feature_layer_norm)

    def _encode(self, *inputs):
        feat, mask, rel, scp = self._sal_features(self.encoder.shared, *inputs)
        enc = self.encoder.encoder(feat, None, compute_dtype=self.compute_dtype, external_rel_bias=rel, scp=scp)
        return enc, mask

    # reference :62-106
    def forward(self, input_ids, src_attention_mask, label_ids, label_attention_mask, tokenized_ocr,
                ocr_attention_mask, ocr_coordinates, ocr_features, tokenized_obj, obj_attention_mask,
                obj_coordinates, obj_features, max_ocr, max_ques):
        enc, mask = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features,
                                 max_ocr, max_ques)
        return self._logits(self.decode(label_ids, enc, mask, label_attention_mask))

    # reference :121-173
    def generate(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques,
                 start_symbol, end_symbol, max_length=20, isgreedy=True, num_beam=2):
        fn = self.greedy_generate if isgreedy else self.beam_generate
        extra = () if isgreedy else (num_beam,)
        return fn(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates, ocr_features,
                  tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr, max_ques, start_symbol,
                  end_symbol, max_length, *extra)

    @torch.no_grad()
    def greedy_generate(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                        ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr,
                        max_ques, start_symbol, end_symbol, max_len=100, use_cache=True):
        enc, mask = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features,
                                 max_ocr, max_ques)
        return self._greedy(enc, mask, start_symbol, end_symbol, max_len, use_cache)

    @torch.no_grad()
    def beam_generate(self, input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                      ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features, max_ocr,
                      max_ques, start_symbol, end_symbol, max_len=100, num_beam=2):
        enc, mask = self._encode(input_ids, src_attention_mask, tokenized_ocr, ocr_attention_mask, ocr_coordinates,
                                 ocr_features, tokenized_obj, obj_attention_mask, obj_coordinates, obj_features,
                                 max_ocr, max_ques)
        return self._beam(enc, mask, start_symbol, end_symbol, max_len, num_beam)
