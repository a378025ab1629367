"""Development aid: torch.profiler kernel table for one training step of the bench workload."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoneme_vqa_b200 import models, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
cfg = synthetic.t5_config("base")
torch.manual_seed(0)
model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev).set_compute_dtype(torch.bfloat16)
model.train()
optim = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.98), eps=1e-9,
                         fused=True)
b = synthetic.phoneme_latr_batch(B, cfg.vocab_size, device=dev)


def step():
    labels = b["label_ids"]
    loss = model.forward_loss(b["pixel_values"], b["coordinates"], b["input_ids"], labels[:, :-1],
                              b["src_attention_mask"], b["label_attention_mask"][:, :-1], b["ocr_attention_mask"],
                              b["tokenized_ocr"], targets=labels[:, 1:], ignore_index=synthetic.PAD_ID)
    optim.zero_grad(set_to_none=True)
    loss.backward()
    optim.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
from collections import defaultdict  # noqa: E402
from torch.autograd import DeviceType  # noqa: E402

agg = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        a = agg[e.name]
        a[0] += 1
        a[1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if "pvqa::" in k)
print(f"GPU busy {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches; libpvqa kernels {mine / 1e3:.3f} ms "
      f"({100 * mine / tot:.1f}%)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:3d}  {name[:110]}")
