"""CPU restatement of the reference MODELS (TEST INFRASTRUCTURE ONLY — see ref_ops.py header).

The reference composes HuggingFace ``T5EncoderModel`` / ``ViTModel`` and
``torch.nn.TransformerDecoder``; those libraries are third-party code that is not under
/root/reference (requirements.txt:1, unpinned; installed here: transformers 5.5.0,
torch 2.11.0).  The restatement therefore composes the *same library modules* in the
same way as core/model/PhonemeLaTr.py does, with ``from_pretrained`` replaced by
config-init (no network) and the 3-table ``PhonemeEmbedding`` the call site expects
(SURVEY.md D1).  It is pinned to the real reference by tests/golden/model_*.npz, written
by oracle/make_golden.py from the reference classes themselves.

Used as: parity checker in tests/, CPU baseline in bench.py (``cpu_baseline`` /
``--impl reference``).
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch
import torch.nn as nn
from transformers import T5Config, T5EncoderModel, T5ForConditionalGeneration, ViTConfig, ViTModel

from . import ref_ops


def make_config(d_model=768, d_kv=64, num_heads=12, d_ff=3072, num_layers=12, vocab_size=36096,
                num_decoder_layers=4, n_head=12, max_2d_position_embeddings=1024, vit_config=None,
                dropout_rate=0.1, feed_forward_proj="relu"):
    cfg = T5Config(d_model=d_model, d_kv=d_kv, num_heads=num_heads, d_ff=d_ff, num_layers=num_layers,
                   vocab_size=vocab_size, dropout_rate=dropout_rate, feed_forward_proj=feed_forward_proj,
                   decoder_start_token_id=0)
    cfg.update({"max_2d_position_embeddings": max_2d_position_embeddings, "vit_model": "random-init",
                "num_decoder_layers": num_decoder_layers, "n_head": n_head, "random_init": True,
                "vit_config": vit_config})
    return cfg


TINY_VIT = dict(hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                image_size=32, patch_size=16)


def tiny_config(**kw):
    base = dict(d_model=192, d_kv=64, num_heads=3, d_ff=256, num_layers=2, vocab_size=120,
                num_decoder_layers=2, n_head=3, vit_config=TINY_VIT)
    base.update(kw)
    return make_config(**base)


# core/model/PhonemeLaTr.py:17-44
class SpatialModule(nn.Module):
    def __init__(self, config):
        super().__init__()
        n, d = config.max_2d_position_embeddings, config.d_model
        self.top_left_x = nn.Embedding(n, d)
        self.bottom_right_x = nn.Embedding(n, d)
        self.top_left_y = nn.Embedding(n, d)
        self.bottom_right_y = nn.Embedding(n, d)
        self.width_emb = nn.Embedding(n, d)
        self.height_emb = nn.Embedding(n, d)

    def forward(self, coordinates):
        return ref_ops.spatial_module(coordinates, [self.top_left_x.weight, self.top_left_y.weight,
                                                    self.bottom_right_x.weight, self.bottom_right_y.weight,
                                                    self.width_emb.weight, self.height_emb.weight])


# PhonoLaTr/modules.py:27-63 as called at core/model/PhonemeLaTr.py:72-78
class PhonemeEmbedding(nn.Module):
    def __init__(self, on_v, rh_v, to_v, on_dim, rt_dim):
        super().__init__()
        self.onset_embedding = nn.Embedding(on_v, on_dim)
        self.rhyme_embedding = nn.Embedding(rh_v, rt_dim)
        self.tone_embedding = nn.Embedding(to_v, rt_dim)

    def forward(self, t):
        return ref_ops.phoneme_embedding(t, self.onset_embedding.weight, self.rhyme_embedding.weight,
                                         self.tone_embedding.weight)


# core/model/modules/transformer_utils.py:6-25
class SinusoidalPositionalEncoding(nn.Module):
    def __init__(self, emb_size, dropout, maxlen=5000):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("pos_embedding", ref_ops.sinusoidal_table(emb_size, maxlen))

    def forward(self, x):
        return self.dropout(ref_ops.positional_encoding(x, self.pos_embedding))


# core/model/modules/transformer_utils.py:38-64
class BaseDecoder(nn.Module):
    def __init__(self, emb_size, num_layers, n_head):
        super().__init__()
        self.decoder = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=emb_size, nhead=n_head, batch_first=True), num_layers=num_layers)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None):
        return self.decoder(tgt=tgt, memory=memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                            tgt_key_padding_mask=tgt_key_padding_mask,
                            memory_key_padding_mask=memory_key_padding_mask)


def _vit_from(config):
    vc = getattr(config, "vit_config", None)
    return ViTModel(ViTConfig(**vc) if isinstance(vc, dict) else ViTConfig())


# core/model/PhonemeLaTr.py:46-236
class PhonemeLaTr(nn.Module):
    def __init__(self, config, onset_vocab_size, rhyme_vocab_size, tone_vocab_size):
        super().__init__()
        self.config = config
        self.encoder = T5EncoderModel(config)
        self.spatial_feat_extractor = SpatialModule(config)
        self.vit = _vit_from(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        for _, child in self.vit.named_children():
            for p in child.parameters():
                p.requires_grad = False
        d = config.d_model
        self.rhyme_tone_embed_dim = d // 3
        self.onset_embed_dim = int(d - self.rhyme_tone_embed_dim * 2)
        self.tgt_tok_emb = PhonemeEmbedding(onset_vocab_size, rhyme_vocab_size, tone_vocab_size,
                                            self.onset_embed_dim, self.rhyme_tone_embed_dim)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(d, config.num_decoder_layers, config.n_head)
        self.shared_lm_head = nn.Linear(d, d)
        self.onset_lm_head = nn.Linear(self.onset_embed_dim, onset_vocab_size)
        self.rhyme_lm_head = nn.Linear(self.rhyme_tone_embed_dim, rhyme_vocab_size)
        self.tone_lm_head = nn.Linear(self.rhyme_tone_embed_dim, tone_vocab_size)

    def _calculate_embedding(self, pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask,
                             tokenized_ocr):
        img_feat = self.visual_projector(self.vit(pixel_values).last_hidden_state)
        spatial = self.spatial_feat_extractor(coordinates)
        layout_feat = self.encoder.shared(tokenized_ocr) + spatial
        feat = torch.cat([img_feat, layout_feat, self.encoder.shared(input_ids)], axis=1)
        mask = torch.cat([torch.ones(img_feat.shape[:2]).to(img_feat.device), ocr_attention_mask,
                          src_attention_mask], axis=1)
        return feat, mask

    @staticmethod
    def _square_mask(sz, device):
        m = (torch.triu(torch.ones((sz, sz), device=device)) == 1).transpose(0, 1)
        return m.float().masked_fill(m == 0, float("-inf")).masked_fill(m == 1, float(0.0))

    def decode(self, labels, enc, enc_mask, label_mask=None):
        emb = self.positional_encoding(self.tgt_tok_emb(labels))
        return self.decoder(emb, enc, tgt_mask=self._square_mask(labels.size(1), labels.device),
                            memory_key_padding_mask=enc_mask, tgt_key_padding_mask=label_mask)

    def _split_heads(self, h):
        on, rt = self.onset_embed_dim, self.rhyme_tone_embed_dim
        return (self.onset_lm_head(h[:, :, :on]), self.rhyme_lm_head(h[:, :, on:on + rt]),
                self.tone_lm_head(h[:, :, on + rt:]))

    def forward(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                ocr_attention_mask, tokenized_ocr):
        emb, mask = self._calculate_embedding(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                              src_attention_mask, tokenized_ocr)
        enc = self.encoder(attention_mask=mask, inputs_embeds=emb).last_hidden_state
        dec = self.decode(labels, enc, mask, label_attention_mask)
        return self._split_heads(self.shared_lm_head(dec))

    @torch.no_grad()
    def greedy_generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask,
                        tokenized_ocr, start_symbol, end_symbol, max_len=100):
        bz = input_ids.size(0)
        emb, mask = self._calculate_embedding(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                              src_attention_mask, tokenized_ocr)
        enc = self.encoder(attention_mask=mask, inputs_embeds=emb).last_hidden_state
        ys = torch.tensor([[[start_symbol, 0, 0]]], dtype=torch.long).repeat(bz, 1, 1)
        for _ in range(max_len):
            out = self.decode(ys, enc, mask)
            on, rh, to = self._split_heads(out)        # reference :195-205: no shared_lm_head here
            nxt = torch.stack([on[:, -1].argmax(-1), rh[:, -1].argmax(-1), to[:, -1].argmax(-1)], dim=-1)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
            if torch.any(ys[:, :, 0] == end_symbol, dim=1).sum() == bz:
                break
        return ys


# core/model/LaTr.py:42-111
class LaTr(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.backbone = T5ForConditionalGeneration(config)
        self.spatial_feat_extractor = SpatialModule(config)
        self.vit = _vit_from(config)
        self.visual_projector = nn.Linear(self.vit.config.hidden_size, config.d_model)
        for _, child in self.vit.named_children():
            for p in child.parameters():
                p.requires_grad = False

    def calculate_embedding(self, pixel_values, coordinates, input_ids, ocr_attention_mask, src_attention_mask,
                            tokenized_ocr):
        img_feat = self.visual_projector(self.vit(pixel_values).last_hidden_state)
        layout_feat = self.backbone.shared(tokenized_ocr) + self.spatial_feat_extractor(coordinates)
        feat = torch.cat([img_feat, layout_feat, self.backbone.shared(input_ids)], axis=1)
        mask = torch.cat([torch.ones(img_feat.shape[:2]).to(img_feat.device), ocr_attention_mask,
                          src_attention_mask], axis=1)
        return feat, mask

    def forward(self, pixel_values, coordinates, input_ids, labels, src_attention_mask, label_attention_mask,
                ocr_attention_mask, tokenized_ocr):
        emb, mask = self.calculate_embedding(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                             src_attention_mask, tokenized_ocr)
        enc = self.backbone.encoder(attention_mask=mask, inputs_embeds=emb).last_hidden_state
        dec = self.backbone.decoder(encoder_hidden_states=enc, inputs_embeds=self.backbone.shared(labels),
                                    attention_mask=label_attention_mask).last_hidden_state
        return self.backbone.lm_head(dec)

    def generate(self, pixel_values, coordinates, input_ids, src_attention_mask, ocr_attention_mask, tokenized_ocr,
                 max_length=20):
        emb, _ = self.calculate_embedding(pixel_values, coordinates, input_ids, ocr_attention_mask,
                                          src_attention_mask, tokenized_ocr)
        return self.backbone.generate(inputs_embeds=emb, max_length=max_length, do_sample=False, num_beams=1)


# core/executor/LaTr_Executor.py:140-163 — one training step's loss (labels from the HF tokenizer, pad id 0)
def latr_loss(model, batch, pad_id=0):
    labels = batch["label_ids"]
    logits = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"], input_ids=batch["input_ids"],
                   labels=labels[:, :-1], src_attention_mask=batch["src_attention_mask"],
                   label_attention_mask=batch["label_attention_mask"][:, :-1],
                   ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    return nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1),
                                       ignore_index=pad_id)


def latr_batch(B, cfg, T=19, L_ocr=12, L_q=6, seed=3, image=32):
    """LaTr labels: "<pad> " + answer tokens + eos, padded with 0; attention mask 1 = valid (int64)."""
    b = synthetic_batch(B, cfg, T=T, L_ocr=L_ocr, L_q=L_q, seed=seed, image=image)
    g = torch.Generator().manual_seed(seed + 1)
    labels = torch.zeros(B, T + 1, dtype=torch.long)
    mask = torch.zeros(B, T + 1, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(2, T, (1,), generator=g))
        labels[i, 1:n] = torch.randint(3, cfg.vocab_size, (n - 1,), generator=g)
        labels[i, n] = 1
        mask[i, : n + 1] = 1
    b["label_ids"], b["label_attention_mask"] = labels, mask
    return b


# ----------------------------------------------------------------------------------
# SaL family.  The reference's T52DStack (a copy of HF-4.x T5Stack.forward with `position_bias` injected,
# core/model/modules/SaL_utils.py:226-500) does not run as-is under transformers 5.5, so the encoder is
# restated as a loop over HF T5Block with the external bias: no attention mask is added when the
# bias is external, and layer 0's own relative_attention_bias stays an unused parameter.  Pinning: the real
# classes DO run through the call-convention adapter of oracle/make_golden_variants.py; their outputs
# (tests/golden/model_{sal,customizedsal,phonemesal}_tiny.npz, sal_bias.npz) are reproduced bit for bit by these
# restatements (tests/test_variants_cpu.py).
# ----------------------------------------------------------------------------------
def scp_distance_lut(grid=11):
    """core/model/modules/SaL_utils.py:171-195: 5 * Euclidean cell distance, (x, y, x', y')."""
    xs, ys = np.mgrid[0:grid, 0:grid]
    out = np.zeros((grid, grid, grid, grid))
    for x in range(grid):
        for y in range(grid):
            out[x, y] = np.sqrt((xs - x) ** 2 + (ys - y) ** 2)
    return out * 5


def sal_position_bias(rel_table, scp_table, S, coordinates, max_ques, max_ocr):
    """core/model/modules/SaL_utils.py:81-120,131-139,152-168,208-223 -> (B,H,S,S)."""
    B = coordinates.shape[0]
    pos = torch.arange(S)
    rel = ref_ops.t5_relative_bucket((pos[None, :] - pos[:, None]).numpy(), True, 32, 128)
    bias = torch.nn.functional.embedding(torch.as_tensor(rel).to(rel_table.device), rel_table).permute(2, 0, 1)[None].repeat(B, 1, 1, 1)
    xc = coordinates[:, :, [0, 2]].mean(dim=-1)
    yc = coordinates[:, :, [1, 3]].mean(dim=-1)
    xi = np.int32(np.floor(xc.cpu().numpy() * 11))
    yi = np.int32(np.floor(yc.cpu().numpy() * 11))
    lut = scp_distance_lut()
    d = lut[xi[:, :, None], yi[:, :, None], xi[:, None, :], yi[:, None, :]]          # (B,L,L)
    bk = ref_ops.t5_relative_bucket(torch.tensor(d).to(torch.long).numpy(), True, 32, 100)
    scp = torch.nn.functional.embedding(torch.as_tensor(bk).to(scp_table.device), scp_table).permute(0, 3, 1, 2)
    bias[:, :, max_ques:max_ques + max_ocr, max_ques:max_ques + max_ocr] += scp
    return bias


class _BiasTable(nn.Module):
    def __init__(self, num_heads):
        super().__init__()
        self.relative_attention_bias = nn.Embedding(32, num_heads)


class _BiasAggregated(nn.Module):
    def __init__(self, num_heads):
        super().__init__()
        self.Relative1D = _BiasTable(num_heads)
        self.SCP = _BiasTable(num_heads)


# core/model/PhonemeSaL.py:28-207
class PhonemeSaL(nn.Module):
    def __init__(self, config, vocab_size, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        from transformers.models.t5.modeling_t5 import T5LayerNorm
        self.config = config
        self.vocab_size = vocab_size
        self.encoder = T5EncoderModel(config)
        self.encoder.resize_token_embeddings(config.new_token_embedding_size)
        self.rel2Dbias = _BiasAggregated(config.num_heads)
        d = config.d_model
        self.obj_dropout = nn.Dropout(obj_dropout)
        self.obj_feature_projector = nn.Linear(config.obj_hidden, d)
        self.obj_bbox_projector = nn.Linear(4, d)
        self.obj_feature_layer_norm = T5LayerNorm(d)
        self.ocr_dropout = nn.Dropout(ocr_dropout)
        self.ocr_feature_projector = nn.Linear(config.ocr_hidden, d)
        self.ocr_bbox_projector = nn.Linear(4, d)
        self.ocr_feature_layer_norm = T5LayerNorm(d)
        self.tgt_tok_emb = nn.Embedding(vocab_size, d)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(d, config.num_decoder_layers, config.n_head)
        self.lm_head = nn.Linear(d, vocab_size)
        self.loss_fn = nn.CrossEntropyLoss(ignore_index=0)

    def _encode(self, b):
        obj = (self.obj_feature_layer_norm(self.obj_feature_projector(b["obj_features"]))
               + self.obj_feature_layer_norm(self.obj_bbox_projector(b["obj_coordinates"]))
               + self.encoder.shared(b["tokenized_obj"]))
        ocr = (self.ocr_feature_layer_norm(self.ocr_feature_projector(b["ocr_features"]))
               + self.ocr_feature_layer_norm(self.ocr_bbox_projector(b["ocr_coordinates"]))
               + self.encoder.shared(b["tokenized_ocr"]))
        feat = torch.cat([self.encoder.shared(b["input_ids"]), ocr, obj], dim=1)
        mask = torch.cat([b["src_attention_mask"], b["ocr_attention_mask"], b["obj_attention_mask"]], dim=1)
        bias = sal_position_bias(self.rel2Dbias.Relative1D.relative_attention_bias.weight,
                                 self.rel2Dbias.SCP.relative_attention_bias.weight, feat.shape[1],
                                 b["ocr_coordinates"], b["max_ques"], b["max_ocr"])
        stack = self.encoder.encoder
        h = stack.dropout(feat)
        for blk in stack.block:
            h = blk(h, attention_mask=None, position_bias=bias)[0]
        h = stack.dropout(stack.final_layer_norm(h))
        return h, mask

    def decode(self, labels, enc, enc_mask, label_mask=None):
        emb = self.positional_encoding(self.tgt_tok_emb(labels))
        return self.decoder(emb, enc, tgt_mask=PhonemeLaTr._square_mask(labels.size(1), labels.device),
                            memory_key_padding_mask=enc_mask, tgt_key_padding_mask=label_mask)

    def forward(self, b):
        enc, mask = self._encode(b)
        dec = self.decode(b["label_ids"], enc, mask, b["label_attention_mask"])
        logits = self.lm_head(dec)
        loss = self.loss_fn(logits.reshape((-1, self.vocab_size)), b["shifted_right_label_ids"].reshape(-1))
        return logits, loss

    @torch.no_grad()
    def generate(self, b, start_symbol, end_symbol, max_len=100):
        enc, mask = self._encode(b)
        bz = b["input_ids"].size(0)
        ys = torch.tensor([start_symbol], dtype=torch.long).repeat(bz, 1)
        brk = torch.zeros_like(ys).fill_(0)
        for _ in range(max_len):
            out = self.decode(ys, enc, mask)
            nxt = torch.argmax(self.lm_head(out)[:, -1], dim=-1)
            brk = torch.where((nxt == end_symbol)[:, None], 1, brk)
            ys = torch.cat([ys, nxt.unsqueeze(1)], dim=1)
            if torch.all(brk):
                break
        return ys


def sal_config(**kw):
    cfg = tiny_config(**kw)
    cfg.update({"ocr_hidden": 24, "obj_hidden": 40, "new_token_embedding_size": 130})
    return cfg


def sal_batch(B, cfg, T=11, L_q=16, L_ocr=32, L_obj=16, vocab=253, seed=5):
    """field names / dtypes of core/data/PhonemeSaLDataset.py:78-92 (float masks, bool label mask, boxes in [0,1))."""
    g = torch.Generator().manual_seed(seed)
    V = cfg.new_token_embedding_size

    def toks(L):
        ids = torch.zeros(B, L, dtype=torch.long)
        mask = torch.zeros(B, L)
        for i in range(B):
            n = int(torch.randint(1, L, (1,), generator=g))
            ids[i, :n] = torch.randint(3, V, (n,), generator=g)
            ids[i, n] = 1
            mask[i, : n + 1] = 1
        return ids, mask

    q, qm = toks(L_q)
    ocr, om = toks(L_ocr)
    obj, bm = toks(L_obj)

    def boxes(L):
        xy = torch.rand(B, L, 2, generator=g) * 0.8
        wh = torch.rand(B, L, 2, generator=g) * 0.19
        return torch.cat([xy, xy + wh], dim=-1)

    labels = torch.zeros(B, T + 1, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(2, T, (1,), generator=g))
        labels[i, 0] = 1
        labels[i, 1:n] = torch.randint(4, vocab, (n - 1,), generator=g)
        labels[i, n] = 2
    return {"input_ids": q, "src_attention_mask": qm, "label_ids": labels[:, :-1],
            "shifted_right_label_ids": labels[:, 1:], "label_attention_mask": labels[:, :-1] == 0,
            "tokenized_ocr": ocr, "ocr_attention_mask": om, "ocr_coordinates": boxes(L_ocr),
            "ocr_features": torch.randn(B, L_ocr, cfg.ocr_hidden, generator=g),
            "tokenized_obj": obj, "obj_attention_mask": bm, "obj_coordinates": boxes(L_obj),
            "obj_features": torch.randn(B, L_obj, cfg.obj_hidden, generator=g), "max_ocr": L_ocr, "max_ques": L_q}


# core/executor/PhonemeLaTr_Executor.py:161-196 — one training step's loss
def phoneme_latr_loss(model, batch, pad_id):
    labels = batch["label_ids"]
    trg_input = labels[:, :-1]
    on, rh, to = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"],
                       input_ids=batch["input_ids"], labels=trg_input,
                       src_attention_mask=batch["src_attention_mask"],
                       label_attention_mask=batch["label_attention_mask"][:, :-1],
                       ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    ce = nn.functional.cross_entropy
    return (ce(on.reshape(-1, on.shape[-1]), labels[:, 1:, 0].reshape(-1), ignore_index=pad_id)
            + ce(rh.reshape(-1, rh.shape[-1]), labels[:, 1:, 1].reshape(-1), ignore_index=pad_id)
            + ce(to.reshape(-1, to.shape[-1]), labels[:, 1:, 2].reshape(-1), ignore_index=pad_id))


# ----------------------------------------------------------------------------------
# deterministic weights / batches shared by the golden generator, the tests and bench.py
# ----------------------------------------------------------------------------------
def deterministic_state_dict(model: nn.Module, scale=0.05) -> dict:
    """Every floating tensor of state_dict() is replaced by a function of (key, shape) only
    (numpy legacy RandomState, stable across versions) so fixtures need not store weights."""
    sd = {}
    for k, v in model.state_dict().items():
        if not v.is_floating_point() or k.endswith("pos_embedding"):
            sd[k] = v.clone()
            continue
        rs = np.random.RandomState(zlib.crc32(k.replace("encoder.encoder.embed_tokens", "encoder.shared").encode()))
        x = rs.standard_normal(tuple(v.shape)).astype(np.float32) * scale
        if "norm" in k and k.endswith("weight"):
            x = 1.0 + x
        sd[k] = torch.from_numpy(x).reshape(v.shape)
    return sd


def synthetic_batch(B, cfg, T=127, L_ocr=100, L_q=30, V_sub=(84, 187, 7), seed=1234, image=224,
                    pad_id=2, bos_id=3, eos_id=4):
    """SURVEY.md §8d synthetic batch; field names/dtypes as core/data/PhonemeLaTrDataset.py:51-58."""
    g = torch.Generator().manual_seed(seed)
    V = cfg.vocab_size
    coords = torch.zeros(B, L_ocr, 6, dtype=torch.long)
    ocr = torch.zeros(B, L_ocr, dtype=torch.long)
    om = torch.zeros(B, L_ocr)
    q = torch.zeros(B, L_q, dtype=torch.long)
    qm = torch.zeros(B, L_q)
    labels = torch.full((B, T + 1, 3), pad_id, dtype=torch.long)
    lmask = torch.ones(B, T + 1)                       # 1.0 = pad (float(create_mask), PhonemeLaTrDataset.py:55)
    for b in range(B):
        n = int(torch.randint(min(20, L_ocr - 1), min(100, L_ocr), (1,), generator=g))
        x0 = torch.randint(0, 901, (n,), generator=g)
        y0 = torch.randint(0, 901, (n,), generator=g)
        w = torch.randint(1, 101, (n,), generator=g)
        h = torch.randint(1, 101, (n,), generator=g)
        coords[b, :n] = torch.stack([x0, y0, x0 + w, y0 + h, w, h], dim=-1)
        coords[b, n] = 1000
        ocr[b, :n] = torch.randint(3, V, (n,), generator=g)
        ocr[b, n] = 1
        om[b, : n + 1] = 1
        nq = int(torch.randint(min(8, L_q), L_q + 1, (1,), generator=g))
        q[b, : nq - 1] = torch.randint(3, V, (nq - 1,), generator=g)
        q[b, nq - 1] = 1
        qm[b, :nq] = 1
        ln = int(torch.randint(min(4, T), min(40, T) + 1, (1,), generator=g))
        labels[b, 0] = torch.tensor([bos_id, 0, 0])
        labels[b, 1:ln, 0] = torch.randint(5, V_sub[0], (ln - 1,), generator=g)
        labels[b, 1:ln, 1] = torch.randint(2, V_sub[1], (ln - 1,), generator=g)
        labels[b, 1:ln, 2] = torch.randint(0, V_sub[2], (ln - 1,), generator=g)
        labels[b, ln] = torch.tensor([eos_id, 0, 0])
        lmask[b, : ln + 1] = 0
    pix = torch.randn(B, 3, image, image, generator=g)
    return {"pixel_values": pix, "coordinates": coords, "input_ids": q, "src_attention_mask": qm,
            "label_ids": labels, "label_attention_mask": lmask, "tokenized_ocr": ocr, "ocr_attention_mask": om}


# ----------------------------------------------------------------------------------
# flat-vocabulary (Customized*) batches and the executors' loss: CrossEntropyLoss(ignore_index=pad) on
# logits[:, :-1] vs labels[:, 1:]  (core/executor/CustomizedLaTr_Executor.py:160-182, PreSTU_Executor.py:135-153)
# ----------------------------------------------------------------------------------
def flat_batch(B, cfg, T=13, L_ocr=12, L_q=6, vocab=50, seed=21, image=32, pad_id=0, bos_id=1, eos_id=2):
    """CustomizedLaTrDataset fields (core/data/CustomizedLaTrDataset.py:40-58): int64 src / ocr masks, flat label
    ids, label mask = (ids == pad) as bool."""
    b = synthetic_batch(B, cfg, T=T, L_ocr=L_ocr, L_q=L_q, seed=seed, image=image)
    g = torch.Generator().manual_seed(seed + 1)
    labels = torch.full((B, T + 1), pad_id, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(2, T, (1,), generator=g))
        labels[i, 0] = bos_id
        labels[i, 1:n] = torch.randint(3, vocab, (n - 1,), generator=g)
        labels[i, n] = eos_id
    b["label_ids"], b["label_attention_mask"] = labels, labels == pad_id
    b["src_attention_mask"] = b["src_attention_mask"].long()
    b["ocr_attention_mask"] = b["ocr_attention_mask"].long()
    return b


LATR_KEYS = ("pixel_values", "coordinates", "input_ids", "src_attention_mask", "ocr_attention_mask", "tokenized_ocr")
PRESTU_KEYS = ("pixel_values", "input_ids", "src_attention_mask")


def flat_loss(model, batch, keys, pad_id=0):
    labels = batch["label_ids"]
    logits = model(labels=labels[:, :-1], label_attention_mask=batch["label_attention_mask"][:, :-1],
                   **{k: batch[k] for k in keys})
    return nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1),
                                       ignore_index=pad_id)


def prestu_batch(B, cfg, T=19, L_q=14, seed=9, image=32):
    """PreSTUDataset fields: question + OCR text packed into input_ids, int64 masks, T5 labels (1 = valid)."""
    b = latr_batch(B, cfg, T=T, L_ocr=4, L_q=L_q, seed=seed, image=image)
    b["src_attention_mask"] = b["src_attention_mask"].long()
    return {k: b[k] for k in PRESTU_KEYS + ("label_ids", "label_attention_mask")}


# ----------------------------------------------------------------------------------
# SaL (core/model/SaL.py:24-140) and CustomizedSaL (core/model/CustomizedSaL.py:29-335) restated like PhonemeSaL
# above: HF modules composed as the reference composes them, the T52DStack encoder as a loop over HF T5Block with the
# external bias.  Pinned to the real classes' outputs and layouts (tests/golden/model_{sal,customizedsal}_tiny.npz,
# sal_family_layouts.npz).
# ----------------------------------------------------------------------------------
def _sal_encoder_inputs(m, shared, b):
    obj = (m.obj_feature_layer_norm(m.obj_feature_projector(b["obj_features"]))
           + m.obj_feature_layer_norm(m.obj_bbox_projector(b["obj_coordinates"])) + shared(b["tokenized_obj"]))
    ocr = (m.ocr_feature_layer_norm(m.ocr_feature_projector(b["ocr_features"]))
           + m.ocr_feature_layer_norm(m.ocr_bbox_projector(b["ocr_coordinates"])) + shared(b["tokenized_ocr"]))
    feat = torch.cat([shared(b["input_ids"]), ocr, obj], dim=1)
    mask = torch.cat([b["src_attention_mask"], b["ocr_attention_mask"], b["obj_attention_mask"]], dim=1)
    bias = sal_position_bias(m.rel2Dbias.Relative1D.relative_attention_bias.weight,
                             m.rel2Dbias.SCP.relative_attention_bias.weight, feat.shape[1],
                             b["ocr_coordinates"], b["max_ques"], b["max_ocr"])
    return feat, mask, bias


def _external_bias_stack(stack, feat, bias):
    h = stack.dropout(feat)
    for blk in stack.block:
        h = blk(h, attention_mask=None, position_bias=bias)[0]
    return stack.dropout(stack.final_layer_norm(h))


def _sal_projectors(m, config, obj_dropout, ocr_dropout):
    from transformers.models.t5.modeling_t5 import T5LayerNorm
    d = config.d_model
    m.rel2Dbias = _BiasAggregated(config.num_heads)
    m.obj_dropout = nn.Dropout(obj_dropout)
    m.obj_feature_projector = nn.Linear(config.obj_hidden, d)
    m.obj_bbox_projector = nn.Linear(4, d)
    m.obj_feature_layer_norm = T5LayerNorm(d)
    m.ocr_dropout = nn.Dropout(ocr_dropout)
    m.ocr_feature_projector = nn.Linear(config.ocr_hidden, d)
    m.ocr_bbox_projector = nn.Linear(4, d)
    m.ocr_feature_layer_norm = T5LayerNorm(d)


class SaL(nn.Module):
    def __init__(self, config, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        self.config = config
        self.backbone = T5ForConditionalGeneration(config)
        self.backbone.resize_token_embeddings(config.new_token_embedding_size)
        _sal_projectors(self, config, obj_dropout, ocr_dropout)

    def _encode(self, b):
        feat, _, bias = _sal_encoder_inputs(self, self.backbone.shared, b)
        return _external_bias_stack(self.backbone.encoder, feat, bias)

    def forward(self, b):
        enc = self._encode(b)
        dec = self.backbone.decoder(encoder_hidden_states=enc, inputs_embeds=self.backbone.shared(b["label_ids"]),
                                    attention_mask=b["label_attention_mask"]).last_hidden_state
        return self.backbone.lm_head(dec)

    @torch.no_grad()
    def generate(self, b, max_length=20):
        """HF greedy search from the encoder output (what `backbone.generate(inputs_embeds=..., position_bias=...)`
        does in the reference: SaL.py:136-140)."""
        from transformers.modeling_outputs import BaseModelOutput
        feat, _, bias = _sal_encoder_inputs(self, self.backbone.shared, b)
        enc = _external_bias_stack(self.backbone.encoder, feat, bias)
        return self.backbone.generate(inputs_embeds=feat, encoder_outputs=BaseModelOutput(last_hidden_state=enc),
                                      max_length=max_length, do_sample=False, num_beams=1)


def sal_t5_batch(B, cfg, T=11, seed=5, **kw):
    """SaLDataset fields (core/data/SaLDataset.py): T5 labels "<pad> answer </s>" with an int64 1 = valid mask."""
    b = sal_batch(B, cfg, T=T, seed=seed, **kw)
    g = torch.Generator().manual_seed(seed + 100)
    labels = torch.zeros(B, T + 1, dtype=torch.long)
    mask = torch.zeros(B, T + 1, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(2, T, (1,), generator=g))
        labels[i, 1:n] = torch.randint(3, cfg.new_token_embedding_size, (n - 1,), generator=g)
        labels[i, n] = 1
        mask[i, : n + 1] = 1
    b.pop("shifted_right_label_ids")
    b["label_ids_full"], b["label_mask_full"] = labels, mask
    b["label_ids"], b["label_attention_mask"] = labels[:, :-1], mask[:, :-1]
    return b


def sal_t5_loss(model, b, pad_id=0, as_kwargs=False):
    """core/executor/SaL_Executor.py:190-221"""
    if as_kwargs:
        logits = model(**{k: v for k, v in b.items() if not k.endswith("_full")})
    else:
        logits = model(b)
    return nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), b["label_ids_full"][:, 1:].reshape(-1),
                                       ignore_index=pad_id)


class _RefTokenEmbedding(nn.Module):
    # core/model/modules/transformer_utils.py:27-36
    def __init__(self, vocab_size, emb_size):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, emb_size)
        self.emb_size = emb_size

    def forward(self, tokens):
        return self.embedding(tokens.long()) * math.sqrt(self.emb_size)


class CustomizedSaL(nn.Module):
    def __init__(self, config, tgt_vocab_size, obj_dropout=0.1, ocr_dropout=0.1):
        super().__init__()
        self.config = config
        self.encoder = T5EncoderModel(config)
        self.encoder.resize_token_embeddings(config.new_token_embedding_size)
        _sal_projectors(self, config, obj_dropout, ocr_dropout)
        d = config.d_model
        self.tgt_tok_emb = _RefTokenEmbedding(tgt_vocab_size, d)
        self.positional_encoding = SinusoidalPositionalEncoding(d, dropout=0.1)
        self.decoder = BaseDecoder(d, config.num_decoder_layers, config.n_head)
        self.lm_head = nn.Linear(d, tgt_vocab_size)

    def _encode(self, b):
        feat, mask, bias = _sal_encoder_inputs(self, self.encoder.shared, b)
        return _external_bias_stack(self.encoder.encoder, feat, bias), mask

    def decode(self, labels, enc, enc_mask, label_mask=None):
        emb = self.positional_encoding(self.tgt_tok_emb(labels))
        return self.decoder(emb, enc, tgt_mask=PhonemeLaTr._square_mask(labels.size(1), labels.device),
                            memory_key_padding_mask=enc_mask, tgt_key_padding_mask=label_mask)

    def forward(self, b):
        enc, mask = self._encode(b)
        return self.lm_head(self.decode(b["label_ids"], enc, mask, b["label_attention_mask"]))

    @torch.no_grad()
    def greedy_generate(self, b, start_symbol, end_symbol, max_len=100):
        enc, mask = self._encode(b)
        bz = b["input_ids"].size(0)
        ys = torch.ones(bz, 1).fill_(start_symbol).type(torch.long)
        for _ in range(max_len):
            prob = self.lm_head(self.decode(ys, enc, mask)[:, -1])
            ys = torch.cat([ys, torch.argmax(prob, dim=-1).view(bz, -1)], dim=1)
            if torch.any(ys == end_symbol, dim=1).sum() == bz:
                break
        return ys


def customized_sal_batch(B, cfg, T=11, vocab=50, seed=5, **kw):
    """CustomizedSaLDataset fields: flat label ids, label mask = pad positions (bool), float encoder masks."""
    b = sal_batch(B, cfg, T=T, vocab=vocab, seed=seed, **kw)
    full = torch.cat([b["label_ids"], b.pop("shifted_right_label_ids")[:, -1:]], dim=1)
    b["label_ids_full"] = full
    b["label_ids"], b["label_attention_mask"] = full[:, :-1], full[:, :-1] == 0
    return b
