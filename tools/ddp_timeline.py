"""Kernel timeline of one data-parallel training step: where the NCCL all-reduces sit against
the backward, what they run next to, and how much of them is exposed after the last compute kernel.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/ddp_timeline.py [--waves K] [--out gpurun_out/ddp_timeline_nN]
Rank 0 profiles 2 graph replays with torch.profiler (CUPTI kernel activity), writes <out>.txt (summary) and
<out>.kernels.json (name, stream, start, duration of every kernel of ONE step)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phoneme_vqa_b200 as pv  # noqa: E402
from phoneme_vqa_b200 import models, ops, parallel, synthetic, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--waves", type=int, default=None, help="attention-backward CTAs per SM (GradReducer default if omitted)")
ap.add_argument("--bucket-mb", type=float, default=32.0)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/ddp_timeline")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synthetic.t5_config("base")
torch.manual_seed(0)
model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev).set_compute_dtype(torch.bfloat16)
model.train()
ops.manual_seed(1234 + rank)
reducer = None
if world > 1:
    reducer = parallel.GradReducer(model, bucket_mb=args.bucket_mb, attn_bwd_waves=args.waves)
    reducer.broadcast_parameters(0)
tr = train.TrainStep(model, reducer, lr=5e-5, betas=(0.9, 0.98), eps=1e-9, warmup_iters=2000, ignore_index=synthetic.PAD_ID)
batches = [synthetic.phoneme_latr_batch(args.batch, cfg.vocab_size, seed=1234 + rank * 1000 + i, device=dev) for i in range(2)]
for i in range(6):
    tr(batches[i % 2])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    tr(batches[i % 2])
e.record()
torch.cuda.synchronize()
ms_step = a.elapsed_time(e) / 10
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        tr(batches[i % 2])
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank == 0:
    ev = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA and x.time_range.end > x.time_range.start]
    ks = sorted(((x.time_range.start, x.time_range.end - x.time_range.start, x.name) for x in ev
                 if not x.name.lower().startswith(("memcpy", "memset"))), key=lambda t: t[0])
    # the second replay: kernels after the largest gap... simpler: split at the midpoint of the kernel count
    half = len(ks) // 2
    step = ks[half:]
    # what sits between the two replays on the device (every CUDA activity, copies and memsets included)
    allev = sorted(((x.time_range.start, x.time_range.end - x.time_range.start, x.name) for x in ev), key=lambda t: t[0])
    end1 = max(s + d for s, d, _ in ks[:half])
    between = [(s - end1, d, n[:60]) for s, d, n in allev if end1 - 1 <= s < step[0][0]]
    inter = f"device idle between the two replays: {(step[0][0] - end1):.1f} us; activities in it: " + \
        ", ".join(f"{n} @+{s:.0f}us ({d:.1f}us)" for s, d, n in between[:12])
    t0 = step[0][0]
    is_nccl = lambda n: "nccl" in n.lower()      # noqa: E731
    comp = [(s - t0, d, n) for s, d, n in step if not is_nccl(n)]
    nccl = [(s - t0, d, n) for s, d, n in step if is_nccl(n)]
    end = max(s + d for s, d, _ in comp + nccl)
    adam = [c for c in comp if "multi_tensor" in c[2] or "adam" in c[2].lower()]
    first_adam = min((c[0] for c in adam), default=end)
    last_bwd = max((c[0] + c[1] for c in comp if c[0] < first_adam), default=0.0)
    lines = [f"world {world}  attn_bwd_waves {reducer.attn_bwd_waves if reducer else 1}  bucket_mb {args.bucket_mb:g}  step {ms_step:.3f} ms (CUDA events, 10 replays)  "
             f"profiled step span {end / 1e3:.3f} ms  kernels {len(comp)} compute + {len(nccl)} nccl",
             inter,
             f"compute kernel time {sum(c[1] for c in comp) / 1e3:.3f} ms   nccl kernel time {sum(c[1] for c in nccl) / 1e3:.3f} ms",
             f"last backward kernel ends at {last_bwd / 1e3:.3f} ms, first optimizer kernel starts at {first_adam / 1e3:.3f} ms "
             f"(exposed wait {max(0.0, first_adam - last_bwd) / 1e3:.3f} ms)", "",
             f"{'#':>3s} {'start ms':>9s} {'dur us':>9s} {'compute busy us':>16s}  concurrent compute kernels (name x count)"]
    for i, (s, d, n) in enumerate(nccl):
        over = {}
        busy = 0.0
        for cs, cd, cn in comp:
            o = min(s + d, cs + cd) - max(s, cs)
            if o > 0:
                busy += o
                key = cn.split("(")[0].replace("void ", "")[:36]
                over[key] = over.get(key, 0) + 1
        top = ", ".join(f"{k} x{v}" for k, v in sorted(over.items(), key=lambda kv: -kv[1])[:4])
        lines.append(f"{i:3d} {s / 1e3:9.3f} {d:9.1f} {busy:16.1f}  {top}")
    # slow-down of the compute kernels that overlap a collective: same kernel name, overlapping vs not
    import collections
    groups = collections.defaultdict(lambda: [[], []])
    for cs, cd, cn in comp:
        ov = any(min(s + d, cs + cd) - max(s, cs) > 0 for s, d, _ in nccl)
        groups[cn.split("(")[0].replace("void ", "")[:48]][1 if ov else 0].append(cd)
    lines += ["", "compute kernels: mean duration (us) alone vs while a collective is running"]
    for k, (alone, ovl) in sorted(groups.items(), key=lambda kv: -sum(kv[1][1])):
        if alone and ovl and sum(ovl) > 50:
            lines.append(f"   {k:48s} alone {sum(alone) / len(alone):8.1f} (n={len(alone):3d})   overlapped {sum(ovl) / len(ovl):8.1f} (n={len(ovl):3d})")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    open(args.out + ".txt", "w").write("\n".join(lines) + "\n")
    json.dump([{"t_us": round(s - t0, 1), "dur_us": round(d, 1), "name": n[:80]} for s, d, n in step], open(args.out + ".kernels.json", "w"))
    print("\n".join(lines[:40]))
tr.close()
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
