"""Run W warm-up + 1 training steps of the bench workload eagerly (for `ncu` launch lists: the last
`launches_per_step` rows of the CSV are one steady-state step)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phoneme_vqa_b200 as pv  # noqa: E402
from phoneme_vqa_b200 import models, synthetic, train  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
cfg = synthetic.t5_config("base")
torch.manual_seed(0)
model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev).set_compute_dtype(torch.bfloat16)
model.train()
tr = train.TrainStep(model, None, use_graph=False)
b = synthetic.phoneme_latr_batch(B, cfg.vocab_size, device=dev)
for _ in range(W):
    tr(b)
torch.cuda.synchronize()
c0 = pv.launch_count()
torch.cuda.profiler.start()          # ncu --profile-from-start off: only this step is listed (all threads)
loss = tr(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss), "pvqa launches in the step", pv.launch_count() - c0)
