"""fp32-mode gradient error of EVERY parameter at T5-base dims against the CPU oracle (diagnostic for
tests/test_train_gpu.py::test_t5_base_dims_forward_loss_backward_match_oracle)."""
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

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cfg = ref_model.make_config(vit_config=dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=2,
                                            intermediate_size=128, image_size=224, patch_size=16),
                            vocab_size=2048, num_layers=layers)
oracle, model = T._pair(cfg)
batch = ref_model.synthetic_batch(2, cfg, T=127, L_ocr=100, L_q=30, V_sub=T.VOCAB, seed=21, image=224)
oracle.train(); model.train()
T._no_dropout(oracle); T._no_dropout(model)
ref_loss = ref_model.phoneme_latr_loss(oracle, batch, 2)
ref_loss.backward()
# the same oracle in float64: how far is the fp32 CPU oracle itself from exact?
o64 = ref_model.PhonemeLaTr(cfg, *T.VOCAB)
o64.load_state_dict(oracle.state_dict())
o64 = o64.double(); o64.train(); T._no_dropout(o64)
b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
l64 = ref_model.phoneme_latr_loss(o64, b64, 2)
l64.backward()
b = T._to(batch, T.DEV)
loss = T._loss(model, b)
loss.backward()
print("loss", loss.item(), "oracle32", ref_loss.item(), "oracle64", l64.item())
ref = dict(oracle.named_parameters()); r64 = dict(o64.named_parameters())
rows = []
for name, p in model.named_parameters():
    if p.grad is None or ref[name].grad is None:
        continue
    e = r64[name].grad
    a, r = p.grad.double().cpu(), ref[name].grad.double()
    n = float(e.norm()) + 1e-30
    rows.append((name, float((a - e).norm()) / n, float((r - e).norm()) / n, float((a - r).norm()) / n))
print(f"{'parameter':80s} {'ours-f64':>10s} {'orc32-f64':>10s} {'ours-orc32':>10s}")
for name, x, y, z in rows:
    if x > 2e-4 or y > 2e-4:
        print(f"{name:80s} {x:10.2e} {y:10.2e} {z:10.2e}")
print("worst ours", max(rows, key=lambda t: t[1])[:2], "worst oracle32", max(rows, key=lambda t: t[2])[::2])
