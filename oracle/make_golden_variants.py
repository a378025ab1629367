"""Goldens for the rest of the model family from the REAL reference classes (run in the build container only):

    python oracle/make_golden_variants.py        # needs /root/reference, CPU only

CustomizedLaTr, CustomizedPreSTU and PreSTU run from the reference as they are once `from_pretrained` is replaced by
config-init (no network); their logits, loss, gradient norms, greedy / beam ids and state_dict layout go to
tests/golden/model_<name>_tiny.npz.  SaL / CustomizedSaL cannot run here (their T52DStack is written against
transformers 4.x — SURVEY D8); for those only the state_dict layout of the constructed reference modules is
recorded when construction succeeds, the numerics are checked against the oracle restatement.
"""
import importlib
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


def _patch(mod, cfg, t5_cls):
    import transformers
    setattr(mod, t5_cls, type("T5", (), {"from_pretrained": staticmethod(lambda name: getattr(transformers, t5_cls)(cfg))}))
    mod.ViTModel = type("ViT", (), {"from_pretrained": staticmethod(lambda name: ref_model._vit_from(cfg))})


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0


def _record(model, loss_fn, out):
    model.train()
    _no_dropout(model)
    loss = loss_fn(model)
    loss.backward()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    out["grad_keys"] = np.array(sorted(grads.keys()))
    out["grad_norms"] = np.array([grads[k].double().norm().item() for k in sorted(grads.keys())])
    out["state_dict_keys"] = np.array(list(model.state_dict().keys()))
    out["state_dict_shapes"] = np.array([json.dumps(list(v.shape)) for v in model.state_dict().values()])
    out["frozen"] = np.array(sorted(k for k, p in model.named_parameters() if not p.requires_grad))
    model.eval()


def golden_customized_latr():
    mod = importlib.import_module("core.model.CustomizedLaTr")
    cfg = ref_model.tiny_config()
    _patch(mod, cfg, "T5EncoderModel")
    torch.manual_seed(0)
    model = mod.CustomizedLaTr(cfg, tgt_vocab_size=50)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    b = ref_model.flat_batch(3, cfg)
    model.eval()
    out = {"logits": model(labels=b["label_ids"][:, :-1], label_attention_mask=b["label_attention_mask"][:, :-1],
                           **{k: b[k] for k in ref_model.LATR_KEYS}).detach().numpy()}
    _record(model, lambda m: ref_model.flat_loss(m, b, ref_model.LATR_KEYS), out)
    args = [b[k] for k in ref_model.LATR_KEYS]
    with torch.no_grad():
        out["greedy_ids"] = model.generate(*args, start_symbol=1, end_symbol=2, max_length=7).numpy()
        for nb in (2, 3):
            out[f"beam{nb}_ids"] = model.generate(*args, start_symbol=1, end_symbol=2, max_length=5, isgreedy=False,
                                                  num_beam=nb).numpy()
        # the scores the beam routine starts from, so the selection rule can be replayed on CPU
        ys = torch.ones(3, 1, dtype=torch.long)
        emb, mask = model._calculate_embedding(b["pixel_values"], b["coordinates"], b["input_ids"],
                                               b["ocr_attention_mask"], b["src_attention_mask"], b["tokenized_ocr"])
        enc = model.encoder(attention_mask=mask, inputs_embeds=emb).last_hidden_state
        out["beam_prob"] = model.lm_head(model.decode(ys, enc, mask)[:, -1]).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_customizedlatr_tiny.npz"), **out)
    print("CustomizedLaTr: loss", float(out["loss"]), "greedy", out["greedy_ids"].tolist(), "beam2", out["beam2_ids"].tolist())


def golden_customized_prestu():
    mod = importlib.import_module("core.model.CustomizedPreSTU")
    cfg = ref_model.tiny_config()
    _patch(mod, cfg, "T5EncoderModel")
    torch.manual_seed(0)
    model = mod.CustomizedPreSTU(cfg, tgt_vocab_size=50)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    b = ref_model.flat_batch(3, cfg, seed=23)
    model.eval()
    out = {"logits": model(labels=b["label_ids"][:, :-1], label_attention_mask=b["label_attention_mask"][:, :-1],
                           **{k: b[k] for k in ref_model.PRESTU_KEYS}).detach().numpy()}
    _record(model, lambda m: ref_model.flat_loss(m, b, ref_model.PRESTU_KEYS), out)
    with torch.no_grad():
        out["greedy_ids"] = model.generate(*[b[k] for k in ref_model.PRESTU_KEYS], start_symbol=1, end_symbol=2,
                                           max_length=7).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_customizedprestu_tiny.npz"), **out)
    print("CustomizedPreSTU: loss", float(out["loss"]), "greedy", out["greedy_ids"].tolist())


def golden_prestu():
    mod = importlib.import_module("core.model.PreSTU")
    cfg = ref_model.tiny_config()
    _patch(mod, cfg, "T5ForConditionalGeneration")
    torch.manual_seed(0)
    model = mod.PreSTU(cfg)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    b = ref_model.prestu_batch(3, cfg)
    model.eval()
    out = {"logits": model(labels=b["label_ids"][:, :-1], label_attention_mask=b["label_attention_mask"][:, :-1],
                           **{k: b[k] for k in ref_model.PRESTU_KEYS}).detach().numpy()}
    _record(model, lambda m: ref_model.flat_loss(m, b, ref_model.PRESTU_KEYS), out)
    with torch.no_grad():
        out["generate_ids"] = model.generate(*[b[k] for k in ref_model.PRESTU_KEYS], max_length=8).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_prestu_tiny.npz"), **out)
    print("PreSTU: loss", float(out["loss"]), "generate", out["generate_ids"].tolist())


def golden_phoneme_prestu():
    """PhonemePreSTU with the two documented shims: the 3-table PhonemeEmbedding and the
    `calculate_embedding` name its forward calls (the class defines `_calculate_embedding` — SURVEY D5).  Its
    `greedy_generate` still has the LaTr argument list and cannot run, so no ids are recorded."""
    from oracle.make_golden import ShimPhonemeEmbedding
    mod = importlib.import_module("core.model.PhonemePreSTU")
    cfg = ref_model.tiny_config()
    _patch(mod, cfg, "T5EncoderModel")
    mod.PhonemeEmbedding = ShimPhonemeEmbedding
    mod.PhonemePreSTU.calculate_embedding = mod.PhonemePreSTU._calculate_embedding
    torch.manual_seed(0)
    vocab = (21, 33, 7)
    model = mod.PhonemePreSTU(cfg, *vocab)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    b = ref_model.synthetic_batch(3, cfg, T=9, L_ocr=4, L_q=14, V_sub=vocab, seed=17, image=32)

    def fwd(m):
        return m(pixel_values=b["pixel_values"], input_ids=b["input_ids"], labels=b["label_ids"][:, :-1],
                 src_attention_mask=b["src_attention_mask"], label_attention_mask=b["label_attention_mask"][:, :-1])

    def loss_fn(m):
        on, rh, to = fwd(m)
        ce, lab = nn.functional.cross_entropy, b["label_ids"]
        return sum(ce(x.reshape(-1, x.shape[-1]), lab[:, 1:, i].reshape(-1), ignore_index=2)
                   for i, x in enumerate((on, rh, to)))

    model.eval()
    on, rh, to = fwd(model)
    out = {"onset_logits": on.detach().numpy(), "rhyme_logits": rh.detach().numpy(), "tone_logits": to.detach().numpy()}
    _record(model, loss_fn, out)
    np.savez_compressed(os.path.join(GOLD, "model_phonemeprestu_tiny.npz"), **out)
    print("PhonemePreSTU: loss", float(out["loss"]), "trainable ViT tensors",
          sum(k.startswith("vit.") for k in out["grad_keys"]))


def golden_sal_layouts():
    """state_dict layout of the reference SaL / CustomizedSaL / PhonemeSaL modules (construction only)."""
    import transformers
    out = {}
    cfg = ref_model.sal_config()
    for name, args in (("SaL", ()), ("CustomizedSaL", (50,)), ("PhonemeSaL", (253,))):
        try:
            mod = importlib.import_module("core.model." + name)
            if name == "SaL":
                # transformers 5.x: T5Stack(config) no longer takes the embedding table; give the reference's
                # `T5Stack(config, shared)` call (SaL_utils.py:513) the 4.x behaviour for construction
                su = importlib.import_module("core.model.modules.SaL_utils")
                hf_stack = transformers.models.t5.modeling_t5.T5Stack

                def stack_4x(config, embed_tokens=None, hf_stack=hf_stack):
                    st = hf_stack(config)
                    if embed_tokens is not None:
                        st.embed_tokens = embed_tokens
                    return st
                su.T5Stack = stack_4x
                base = mod.T52dForConditionalGeneration
                mod.T52dForConditionalGeneration = type("T5", (), {"from_pretrained": staticmethod(lambda n, base=base: base(cfg))})
            else:
                mod.T52DEncoderModel = type("T5", (), {"from_pretrained": staticmethod(
                    lambda n, base=mod.T52DEncoderModel: base(cfg))})
            import functools
            for cls in ("RelativePositionBias1D", "SCPRelativePositionBias"):      # default device is "cuda"
                setattr(mod, cls, functools.partial(getattr(mod, cls), device="cpu"))
            model = getattr(mod, name)(cfg, *args)
            sd = model.state_dict()
            out[name + "_keys"] = np.array(list(sd.keys()))
            out[name + "_shapes"] = np.array([json.dumps(list(v.shape)) for v in sd.values()])
            print(name, "constructed:", len(sd), "tensors")
        except Exception as e:                                       # noqa: BLE001
            print(name, "cannot be constructed here:", type(e).__name__, e)
    if out:
        np.savez_compressed(os.path.join(GOLD, "sal_family_layouts.npz"), **out)


if __name__ == "__main__":
    golden_customized_latr()
    golden_customized_prestu()
    golden_prestu()
    golden_phoneme_prestu()
    golden_sal_layouts()
