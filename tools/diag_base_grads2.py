"""Where does the 2.5e-3 fp32-mode gradient error below decoder layer 3's linear2 come from?  Hypothesis: ReLU gates
that sit within forward rounding (1e-4 at these dims) of zero flip between implementations; each flip moves the
gradient by one full-size entry.  Measures (a) our pre-activations against the float64 oracle's, flips in valid rows,
(b) the same gradient error with ops.relu_dropout replaced by torch.relu (autograd), (c) the gradient error of the
float64 oracle evaluated with OUR ReLU gates imposed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_train_gpu as T  # noqa: E402
from oracle import ref_model  # noqa: E402
from phoneme_vqa_b200 import ops  # noqa: E402

cfg = ref_model.make_config(vit_config=dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=2,
                                            intermediate_size=128, image_size=224, patch_size=16), vocab_size=2048)
oracle, model = T._pair(cfg)
batch = ref_model.synthetic_batch(2, cfg, T=127, L_ocr=100, L_q=30, V_sub=T.VOCAB, seed=21, image=224)
model.train(); T._no_dropout(model)
o64 = ref_model.PhonemeLaTr(cfg, *T.VOCAB)
o64.load_state_dict(oracle.state_dict())
o64 = o64.double(); o64.train(); T._no_dropout(o64)
b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}

pre64 = {}
for li, layer in enumerate(o64.decoder.decoder.layers):
    layer.linear1.register_forward_hook(lambda m, i, o, li=li: pre64.__setitem__(li, o.detach()))
l64 = ref_model.phoneme_latr_loss(o64, b64, 2)
l64.backward()
exact = {n: p.grad.clone() for n, p in o64.named_parameters() if p.grad is not None}

NAMES = ("decoder.decoder.layers.3.linear2.weight", "decoder.decoder.layers.3.linear1.bias",
         "decoder.decoder.layers.3.linear1.weight", "decoder.decoder.layers.0.linear1.weight",
         "tgt_tok_emb.rhyme_embedding.weight", "encoder.encoder.block.11.layer.1.DenseReluDense.wo.weight")


def report(tag, grads):
    print(tag)
    for n in NAMES:
        e = exact[n]
        print(f"   {n:64s} {float((grads[n].double().cpu() - e).norm() / e.norm()):.2e}")


b = T._to(batch, T.DEV)
seen = []
orig = ops.relu_dropout


def spy(x, p, training):
    seen.append(x.detach())
    return orig(x, p, training)


ops.relu_dropout = spy
import phoneme_vqa_b200.modules as Mod  # noqa: E402
loss = T._loss(model, b)
loss.backward()
report(f"A. product fp32 mode (loss {loss.item():.7f} vs exact {l64.item():.7f})", {n: p.grad for n, p in model.named_parameters() if p.grad is not None})
dec_pre = [t for t in seen if t.shape[-1] == 2048 and t.shape[-2] == 127 or (t.dim() == 2 and t.shape == (2 * 127, 2048))]
print("captured relu inputs:", [tuple(t.shape) for t in seen][-6:])
valid = (batch["label_attention_mask"][:, :-1] > 0)
for li, t in enumerate(dec_pre[-4:]):
    a = t.reshape(2, 127, 2048).double().cpu(); e = pre64[li]
    flips = (a > 0) != (e > 0)
    print(f"   decoder layer {li}: max |pre - exact| {float((a - e).abs().max()):.2e}; gate flips {int(flips.sum())} "
          f"(in rows with a target: {int(flips[valid].sum())}) of {flips.numel()}")
gates = [(t.reshape(2, 127, 2048) > 0).cpu() for t in dec_pre[-4:]]

# B. torch.relu instead of the kernel
ops.relu_dropout = lambda x, p, training: torch.relu(x)
model.zero_grad(set_to_none=True)
T._loss(model, b).backward()
report("B. relu_dropout kernel replaced by torch.relu", {n: p.grad for n, p in model.named_parameters() if p.grad is not None})
ops.relu_dropout = orig

# C. float64 oracle with OUR gates imposed on the four decoder FFNs: what the exact gradient is for our gates
o64.zero_grad(set_to_none=True)
for li, layer in enumerate(o64.decoder.decoder.layers):
    g = gates[li].double()
    layer.activation = (lambda g: (lambda x: x * g.reshape(x.shape)))(g)
lC = ref_model.phoneme_latr_loss(o64, b64, 2)
lC.backward()
exactC = {n: p.grad.clone() for n, p in o64.named_parameters() if p.grad is not None}
print(f"C. float64 oracle with the product's decoder ReLU gates imposed (loss {lC.item():.7f}); product vs that:")
grads = None
model.zero_grad(set_to_none=True)
T._loss(model, b).backward()
for n in NAMES:
    e = exactC[n]
    print(f"   {n:64s} {float((dict(model.named_parameters())[n].grad.double().cpu() - e).norm() / e.norm()):.2e}")
