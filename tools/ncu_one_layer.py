"""One eager training step of a ONE-LAYER PhonemeLaTr at the bench shape (B = 64, d = 768, H = 12, S = 327, T = 127):
every kernel family of libpvqa_sm100.so launches a handful of times at exactly the shapes of the full step, which is
what an `ncu --set full -k regex:pvqa` capture needs (the 12-layer step has ~3 000 launches, ~40 replays each).
    python tools/ncu_one_layer.py          # plain run first, then the same command under ncu (tools/r02_ncu.sh)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phoneme_vqa_b200 as pv  # noqa: E402
from phoneme_vqa_b200 import models, synthetic, train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
cfg = synthetic.t5_config("base", num_layers=1, num_decoder_layers=1,
                          vit_config=dict(hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                          intermediate_size=3072, image_size=224, patch_size=16))
torch.manual_seed(0)
model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev).set_compute_dtype(torch.bfloat16)
model.train()
tr = train.TrainStep(model, None, use_graph=False)
b = synthetic.phoneme_latr_batch(B, cfg.vocab_size, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(2):
    tr(b)
torch.cuda.synchronize()
flush.zero_()                       # the profiled step starts with a cold L2, like a step of the 12-layer model
c0 = pv.launch_count()
torch.cuda.profiler.start()
loss = tr(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss), "pvqa launches in the step", pv.launch_count() - c0)
