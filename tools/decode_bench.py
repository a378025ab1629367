"""Eval-side timing: greedy phoneme decoding of PhonemeLaTr-base at the bench shapes, with the
key/value cache (`BaseDecoder.step`) and as the reference does it (re-running the 4-layer decoder over the growing
prefix, core/model/PhonemeLaTr.py:193-215).  Random weights never emit <eos>, so both arms run exactly `max_len`
steps.  usage: python tools/decode_bench.py [B] [max_len]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phoneme_vqa_b200 import models, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
cfg = synthetic.t5_config("base")
torch.manual_seed(0)
model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev).set_compute_dtype(torch.bfloat16).eval()
b = synthetic.phoneme_latr_batch(B, cfg.vocab_size, device=dev)
args = (b["pixel_values"], b["coordinates"], b["input_ids"], b["src_attention_mask"], b["ocr_attention_mask"],
        b["tokenized_ocr"], synthetic.BOS_ID, -1)
out = {"B": B, "max_len": L}
for name, use_cache, use_graph in (("kv_cache_graphs", True, True), ("kv_cache", True, False), ("reference_loop", False, False)):
    for _ in range(2):
        ys = model.greedy_generate(*args, max_len=L, use_cache=use_cache, use_graph=use_graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ys = model.greedy_generate(*args, max_len=L, use_cache=use_cache, use_graph=use_graph)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[name] = {"ms_per_batch": ms, "samples_per_s": B / ms * 1e3, "tokens": int(ys.shape[1] - 1)}
    out[name + "_ids_checksum"] = int(ys.sum())
print(json.dumps(out, indent=1))
