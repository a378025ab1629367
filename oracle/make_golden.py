"""Generate tests/golden/* from the REAL reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only), CPU only

The reference ships no tests / golden vectors, so the oracle (oracle/ref_ops.py,
oracle/ref_model.py, oracle/ref_text.py) is pinned to outputs of the reference's own
classes and functions imported from /root/reference.  Two documented shims are needed to
run the mid-refactor snapshot:
  * ``from_pretrained`` -> config-init (no network / HF cache here);
  * the 3-table ``PhonemeEmbedding(on_v, rh_v, to_v, on_dim, rt_dim)`` that
    core/model/PhonemeLaTr.py:72-78 calls but the snapshot does not define, written as
    PhonoLaTr/modules.py:40-63 intends (onset/rhyme/tone tables, concat).
Nothing here is imported at test time; the fixtures it writes are.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PVQA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_model  # noqa: E402


class ShimPhonemeEmbedding(nn.Module):
    """Reconstruction of the intended 3-table embedding (see module docstring)."""

    def __init__(self, on_v, rh_v, to_v, on_dim, rt_dim):
        super().__init__()
        self.onset_embedding = nn.Embedding(on_v, on_dim)
        self.rhyme_embedding = nn.Embedding(rh_v, rt_dim)
        self.tone_embedding = nn.Embedding(to_v, rt_dim)

    def forward(self, t):
        return torch.cat((self.onset_embedding(t[:, :, 0]), self.rhyme_embedding(t[:, :, 1]),
                          self.tone_embedding(t[:, :, 2])), dim=-1)


def reference_phoneme_latr(cfg, vocab):
    import transformers
    import importlib
    ref_mod = importlib.import_module('core.model.PhonemeLaTr')

    ref_mod.PhonemeEmbedding = ShimPhonemeEmbedding
    ref_mod.T5EncoderModel = type("T5Enc", (), {"from_pretrained": staticmethod(lambda name: transformers.T5EncoderModel(cfg))})
    ref_mod.ViTModel = type("ViT", (), {"from_pretrained": staticmethod(lambda name: ref_model._vit_from(cfg))})
    return ref_mod.PhonemeLaTr(cfg, *vocab)


def golden_model():
    torch.manual_seed(0)
    cfg = ref_model.tiny_config()
    vocab = (21, 33, 7)
    model = reference_phoneme_latr(cfg, vocab)
    sd = ref_model.deterministic_state_dict(model)
    model.load_state_dict(sd, strict=True)
    batch = ref_model.synthetic_batch(3, cfg, T=9, L_ocr=12, L_q=6, V_sub=vocab, seed=7, image=32)
    # make the key-padding masks bite: some encoder keys padded, some label pads
    out = {}
    model.eval()
    labels = batch["label_ids"]
    on, rh, to = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"],
                       input_ids=batch["input_ids"], labels=labels[:, :-1],
                       src_attention_mask=batch["src_attention_mask"],
                       label_attention_mask=batch["label_attention_mask"][:, :-1],
                       ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    out["onset_logits"], out["rhyme_logits"], out["tone_logits"] = on.detach().numpy(), rh.detach().numpy(), to.detach().numpy()
    # training-mode loss/grads with dropout disabled (p=0) — RNG streams cannot be matched
    model.train()
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0
    pad_id = 2
    loss = ref_model.phoneme_latr_loss(model, batch, pad_id)
    loss.backward()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    out["grad_keys"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    for k in ["onset_lm_head.weight", "tgt_tok_emb.rhyme_embedding.weight", "spatial_feat_extractor.width_emb.weight",
              "encoder.encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              "decoder.decoder.layers.0.self_attn.in_proj_bias", "shared_lm_head.bias"]:
        out["grad::" + k] = grads[k].numpy()
    nz = grads["encoder.shared.weight"].abs().sum(1).nonzero().flatten()
    out["grad_shared_rows"] = nz.numpy()
    out["grad_shared_vals"] = grads["encoder.shared.weight"][nz].numpy()
    model.eval()
    ys = model.greedy_generate(batch["pixel_values"], batch["coordinates"], batch["input_ids"],
                               batch["src_attention_mask"], batch["ocr_attention_mask"], batch["tokenized_ocr"],
                               start_symbol=3, end_symbol=4, max_len=6)
    out["greedy_ids"] = ys.numpy()
    out["state_dict_keys"] = np.array(list(model.state_dict().keys()))
    out["state_dict_shapes"] = np.array([json.dumps(list(v.shape)) for v in model.state_dict().values()])
    np.savez_compressed(os.path.join(GOLD, "model_phonemelatr_tiny.npz"), **out)
    print("model golden: loss", loss.item(), "keys", len(out["state_dict_keys"]))


def golden_latr():
    import importlib
    import transformers
    ref_mod = importlib.import_module('core.model.LaTr')
    cfg = ref_model.tiny_config()
    ref_mod.T5ForConditionalGeneration = type("T5", (), {"from_pretrained": staticmethod(
        lambda name: transformers.T5ForConditionalGeneration(cfg))})
    ref_mod.ViTModel = type("ViT", (), {"from_pretrained": staticmethod(lambda name: ref_model._vit_from(cfg))})
    torch.manual_seed(0)
    model = ref_mod.LaTr(cfg)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    batch = ref_model.latr_batch(3, cfg)
    model.eval()
    labels = batch["label_ids"]
    logits = model(pixel_values=batch["pixel_values"], coordinates=batch["coordinates"], input_ids=batch["input_ids"],
                   labels=labels[:, :-1], src_attention_mask=batch["src_attention_mask"],
                   label_attention_mask=batch["label_attention_mask"][:, :-1],
                   ocr_attention_mask=batch["ocr_attention_mask"], tokenized_ocr=batch["tokenized_ocr"])
    out = {"logits": logits.detach().numpy()}
    model.train()
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0
    loss = ref_model.latr_loss(model, batch)
    loss.backward()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    out["grad_keys"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    model.eval()
    ys = model.generate(batch["pixel_values"], batch["coordinates"], batch["input_ids"], batch["src_attention_mask"],
                        batch["ocr_attention_mask"], batch["tokenized_ocr"], max_length=8)
    out["generate_ids"] = ys.numpy()
    out["state_dict_keys"] = np.array(list(model.state_dict().keys()))
    out["state_dict_shapes"] = np.array([json.dumps(list(v.shape)) for v in model.state_dict().values()])
    np.savez_compressed(os.path.join(GOLD, "model_latr_tiny.npz"), **out)
    print("latr golden: loss", loss.item(), "generate", ys.tolist())


def golden_sal_bias():
    """real outputs of the reference's bias modules (they import; only T52DStack does not run — SURVEY D8)."""
    from core.model.modules.SaL_utils import (RelativePositionBias1D, RelativePositionBiasAggregated,
                                              SCPRelativePositionBias)
    H, B, S, q0, L = 3, 2, 64, 16, 32
    agg = RelativePositionBiasAggregated(Relative1D=RelativePositionBias1D(num_heads=H, device="cpu"),
                                         SCP=SCPRelativePositionBias(num_heads=H, device="cpu"))
    agg.load_state_dict(ref_model.deterministic_state_dict(agg, scale=1.0))
    g = torch.Generator().manual_seed(12)
    xy = torch.rand(B, L, 2, generator=g) * 0.8
    coords = torch.cat([xy, xy + torch.rand(B, L, 2, generator=g) * 0.19], dim=-1)
    out = agg(torch.zeros(B, S, 8), torch.ones(B, S), coords, q0, L)
    np.savez_compressed(os.path.join(GOLD, "sal_bias.npz"), coords=coords.numpy(), bias=out.detach().numpy(),
                        H=H, S=S, q0=q0, L=L,
                        rel_table=agg.Relative1D.relative_attention_bias.weight.detach().numpy(),
                        scp_table=agg.SCP.relative_attention_bias.weight.detach().numpy(),
                        state_dict_keys=np.array(list(agg.state_dict().keys())))
    print("sal bias golden", tuple(out.shape))


def golden_ops():
    """op-level goldens from reference modules: SpatialModule, SinusoidalPositionalEncoding."""
    import importlib
    ref_mod = importlib.import_module('core.model.PhonemeLaTr')
    from core.model.modules.transformer_utils import SinusoidalPositionalEncoding as RefPE

    cfg = ref_model.tiny_config(d_model=48, d_kv=16, num_heads=3)
    g = torch.Generator().manual_seed(3)
    sm = ref_mod.SpatialModule(cfg)
    sd = ref_model.deterministic_state_dict(sm, scale=1.0)
    sm.load_state_dict(sd)
    coords = torch.randint(0, 1001, (2, 5, 6), generator=g)
    out = {"spatial_coords": coords.numpy(), "spatial_out": sm(coords).detach().numpy()}
    pe = RefPE(48, dropout=0.0, maxlen=64)
    x = torch.randn(2, 7, 48, generator=g)
    out["pe_table"] = pe.pos_embedding.numpy()
    out["pe_in"] = x.numpy()
    out["pe_out"] = pe(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "ops_small.npz"), **out)
    print("ops golden written")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    golden_ops()
    golden_model()
    golden_latr()
    golden_sal_bias()
    try:
        from oracle import make_golden_text
        make_golden_text.main()
    except ImportError:
        print("text goldens: generator not present yet")
