"""Goldens for the rest of the model family from the REAL reference classes (run in the build container only):

    python oracle/make_golden_variants.py        # needs /root/reference, CPU only

CustomizedLaTr, CustomizedPreSTU and PreSTU run from the reference as they are once `from_pretrained` is replaced by
config-init (no network); their logits, loss, gradient norms, greedy / beam ids and state_dict layout go to
tests/golden/model_<name>_tiny.npz.

The SaL family (SaL, CustomizedSaL, PhonemeSaL) is written against transformers 4.x: its `T52DStack`
(core/model/modules/SaL_utils.py:226-500, a copy of the 4.x `T5Stack.forward` with `position_bias` injected) calls
`self.get_head_mask` (gone in 5.x) and invokes `T5Block` with 4.x keyword names.  `adapt_t52d_stack` below bridges
exactly those two call conventions — `get_head_mask -> [None] * n`, and a per-block `forward` that drops the 4.x-only
keywords and passes the rest through — so that the reference's OWN classes run: its embedding composition, its bias
modules, its `T52DStack.forward` loop, its decoders, heads and losses.  Their outputs go to
tests/golden/model_{sal,customizedsal,phonemesal}_tiny.npz and pin the oracle restatement (oracle/ref_model.py), which
reproduces them bit for bit.
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


def adapt_t52d_stack(stack):
    """transformers 4.x -> 5.x call conventions for the reference's T52DStack (see module docstring)."""
    stack.get_head_mask = lambda head_mask, n, *a, **k: [None] * n
    for blk in stack.block:
        def fwd(hidden_states, attention_mask=None, position_bias=None, encoder_hidden_states=None,
                encoder_attention_mask=None, encoder_decoder_position_bias=None, layer_head_mask=None,
                cross_attn_layer_head_mask=None, past_key_value=None, use_cache=False, output_attentions=False,
                hf_forward=blk.forward):
            # 5.x returns (hidden, position_bias); the stack re-inserts the key/value slot itself when use_cache is False
            return hf_forward(hidden_states, attention_mask=attention_mask, position_bias=position_bias,
                              encoder_hidden_states=encoder_hidden_states,
                              encoder_attention_mask=encoder_attention_mask,
                              encoder_decoder_position_bias=encoder_decoder_position_bias, use_cache=False,
                              output_attentions=False)
        blk.forward = fwd


def _build_sal_reference(name, cfg, args):
    import functools
    import transformers
    mod = importlib.import_module("core.model." + name)
    if name == "SaL":
        su = importlib.import_module("core.model.modules.SaL_utils")
        hf_stack = transformers.models.t5.modeling_t5.T5Stack

        def stack_4x(config, embed_tokens=None, hf_stack=hf_stack):     # 4.x: T5Stack(config, shared)
            st = hf_stack(config)
            if embed_tokens is not None:
                st.embed_tokens = embed_tokens
            return st
        su.T5Stack = stack_4x
        if not hasattr(mod, "_real_t52d"):
            mod._real_t52d = mod.T52dForConditionalGeneration
        mod.T52dForConditionalGeneration = type("T5", (), {"from_pretrained": staticmethod(lambda n: mod._real_t52d(cfg))})
    else:
        if not hasattr(mod, "_real_t52d"):
            mod._real_t52d = mod.T52DEncoderModel
        mod.T52DEncoderModel = type("T5", (), {"from_pretrained": staticmethod(lambda n: mod._real_t52d(cfg))})
    for cls in ("RelativePositionBias1D", "SCPRelativePositionBias"):      # default device is "cuda"
        real = getattr(mod, cls)
        real = getattr(real, "func", real)
        setattr(mod, cls, functools.partial(real, device="cpu"))
    torch.manual_seed(0)
    model = getattr(mod, name)(cfg, *args)
    adapt_t52d_stack(model.backbone.encoder if name == "SaL" else model.encoder.encoder)
    model.load_state_dict(ref_model.deterministic_state_dict(model), strict=True)
    return model


SAL_KEYS = ("input_ids", "src_attention_mask", "tokenized_ocr", "ocr_attention_mask", "ocr_coordinates", "ocr_features",
            "tokenized_obj", "obj_attention_mask", "obj_coordinates", "obj_features", "max_ocr", "max_ques")


def golden_sal_family():
    """logits / loss / gradient norms / greedy ids of the REAL reference SaL-family classes (adapter above)."""
    # PhonemeSaL: (logits, loss) from the model itself
    cfg = ref_model.sal_config()
    model = _build_sal_reference("PhonemeSaL", cfg, (253,))
    b = ref_model.sal_batch(3, ref_model.sal_config())
    fkeys = ("input_ids", "src_attention_mask", "label_ids", "shifted_right_label_ids", "label_attention_mask") + SAL_KEYS[2:]
    model.eval()
    out = {"logits": model(**{k: b[k] for k in fkeys})[0].detach().numpy()}
    _record(model, lambda m: m(**{k: b[k] for k in fkeys})[1], out)
    with torch.no_grad():
        out["greedy_ids"] = model.generate(*[b[k] for k in SAL_KEYS], 1, 2, max_len=5).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_phonemesal_tiny.npz"), **out)
    print("PhonemeSaL: loss", float(out["loss"]), "greedy", out["greedy_ids"].tolist())

    # CustomizedSaL: logits; the executor owns the loss (CustomizedSaL_Executor.py:255)
    cfg = ref_model.sal_config()
    model = _build_sal_reference("CustomizedSaL", cfg, (50,))
    b = ref_model.customized_sal_batch(3, ref_model.sal_config())
    fkeys = ("input_ids", "src_attention_mask", "label_ids", "label_attention_mask") + SAL_KEYS[2:]
    model.eval()
    out = {"logits": model(**{k: b[k] for k in fkeys}).detach().numpy()}
    _record(model, lambda m: ref_model.sal_t5_loss(m, b, as_kwargs=True), out)
    with torch.no_grad():
        out["greedy_ids"] = model.generate(*[b[k] for k in SAL_KEYS], start_symbol=1, end_symbol=2, max_length=6).numpy()
        out["beam2_ids"] = model.generate(*[b[k] for k in SAL_KEYS], start_symbol=1, end_symbol=2, max_length=4,
                                          isgreedy=False, num_beam=2).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_customizedsal_tiny.npz"), **out)
    print("CustomizedSaL: loss", float(out["loss"]), "greedy", out["greedy_ids"].tolist(), "beam2", out["beam2_ids"].tolist())

    # SaL: logits over the resized T5 vocabulary; loss in the executor (SaL_Executor.py:221)
    cfg = ref_model.sal_config()
    model = _build_sal_reference("SaL", cfg, ())
    b = ref_model.sal_t5_batch(3, ref_model.sal_config())
    model.eval()
    out = {"logits": model(**{k: b[k] for k in fkeys}).detach().numpy()}
    _record(model, lambda m: ref_model.sal_t5_loss(m, b, as_kwargs=True), out)
    try:
        with torch.no_grad():
            out["generate_ids"] = model.generate(*[b[k] for k in SAL_KEYS], max_length=6).numpy()
    except Exception as e:                                           # noqa: BLE001
        print("SaL.generate cannot run under transformers", __import__("transformers").__version__, "->", type(e).__name__,
              str(e)[:120])
    np.savez_compressed(os.path.join(GOLD, "model_sal_tiny.npz"), **out)
    print("SaL: loss", float(out["loss"]), "generate", out.get("generate_ids", np.zeros(0)).tolist())


def golden_sal_layouts():
    """state_dict layout of the reference SaL / CustomizedSaL / PhonemeSaL modules."""
    out = {}
    for name, args in (("SaL", ()), ("CustomizedSaL", (50,)), ("PhonemeSaL", (253,))):
        sd = _build_sal_reference(name, ref_model.sal_config(), args).state_dict()
        out[name + "_keys"] = np.array(list(sd.keys()))
        out[name + "_shapes"] = np.array([json.dumps(list(v.shape)) for v in sd.values()])
        print(name, "constructed:", len(sd), "tensors")
    np.savez_compressed(os.path.join(GOLD, "sal_family_layouts.npz"), **out)


if __name__ == "__main__":
    golden_customized_latr()
    golden_customized_prestu()
    golden_prestu()
    golden_phoneme_prestu()
    golden_sal_layouts()
    golden_sal_family()
