#!/usr/bin/env python
"""bench.py — train samples/s of PhonoLaTr-base on N B200s (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = the reference's `_train_epoch` body (core/executor/PhonemeLaTr_Executor.py:161-198)
on one synthetic batch: forward, 3x cross-entropy, zero_grad, backward, Adam step, LinearLR step.
Workload: PhonemeLaTr with T5-base dims (d 768, 12 L, 12 H), ViT-B/16-224 frozen, 4-layer target
decoder, phoneme vocab (84,187,7), per-GPU batch 64, S = 197+100+30 = 327, T = 127, bf16 compute
with fp32 master weights / residual stream; random-init weights, synthetic data (no network).

  value  : whole-job samples/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e    : same step driven from pinned HOST batches: H2D of every field + loss.item() per step
  roofline / kernels : per-kernel CUDA-event durations measured live in a separate profiled pass
           of the same steps (events around every C-ABI launch), against MEASURED_PEAKS.json
  cpu_baseline : the oracle port of the reference model on the host cores, bounded sample
  --impl reference : the reference arm = that same CPU port, K steps (rank 0 only)
  --impl eager-gpu : the same port run eagerly by PyTorch on one B200 (fp32 and bf16 autocast) — "what you get today"
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "PhonemeLaTr T5-base, per-GPU batch {B}, S=327 (197 ViT + 100 OCR + 30 question), T=127, phoneme vocab 84/187/7"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region (10 steps are ~0.25 s)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference model on the host cores
# --------------------------------------------------------------------------------------
def cpu_reference_run(batch_size, steps, warmup, threads=None):
    from oracle import ref_model
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = ref_model.make_config()                      # T5-base dims, ViT-B/16, 4-layer decoder
    torch.manual_seed(0)
    model = ref_model.PhonemeLaTr(cfg, 84, 187, 7)
    model.train()
    optim = torch.optim.Adam(model.parameters(), lr=5e-5, betas=(0.9, 0.98), eps=1e-9)
    sched = torch.optim.lr_scheduler.LinearLR(optim, total_iters=2000)
    batch = ref_model.synthetic_batch(batch_size, cfg, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss = ref_model.phoneme_latr_loss(model, batch, 2)
        optim.zero_grad()
        loss.backward()
        optim.step()
        sched.step()
        _ = loss.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch_size * len(times) / total, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} fwd+loss+bwd+Adam step(s) of the oracle port (HF T5/ViT + nn.TransformerDecoder, "
                      f"fp32) at batch {batch_size} of the same PhonoLaTr-base workload, after {warmup} warm-up",
            "ms_per_step": 1e3 * total / len(times)}


def run_eager_gpu_arm(args):
    """SURVEY §8d's "what you get today" number: the oracle port of the reference model (HF T5 / ViT modules +
    nn.TransformerDecoder, i.e. the reference's own module graph) run eagerly by PyTorch on ONE B200, fp32 and
    under torch.autocast(bfloat16), same batch shape as the B200 arm.  Not part of the driver's contract (it runs
    `--impl b200` and `--impl reference`); kept so the comparison can be re-measured with one command."""
    from oracle import ref_model
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.manual_seed(0)
    metric = "train samples/sec (PhonoLaTr-base)"
    workload = WORKLOAD.format(B=args.batch)
    if args.workload == "latr":              # BASELINE config 2: HF T5ForConditionalGeneration + 36 096-way head
        cfg = ref_model.make_config(num_decoder_layers=12)
        model = ref_model.LaTr(cfg).to(dev).train()
        batch = ref_model.latr_batch(args.batch, cfg, T=127, L_ocr=100, L_q=30, seed=1234, image=224)
        loss_of = lambda: ref_model.latr_loss(model, batch, 0)      # noqa: E731
        metric = "train samples/sec (LaTr-base)"
        workload = f"LaTr T5-base (12+12 layers, 36096-way vocabulary head), per-GPU batch {args.batch}, S=327, T=127"
    elif args.workload == "phonosal":        # BASELINE config 5: T5-large dims, S = 464
        cfg = ref_model.make_config(d_model=1024, num_heads=16, d_ff=4096, num_layers=24, n_head=16)
        cfg.update({"ocr_hidden": 512, "obj_hidden": 2048, "new_token_embedding_size": cfg.vocab_size})
        model = ref_model.PhonemeSaL(cfg, 253).to(dev).train()
        batch = ref_model.sal_batch(args.batch, cfg, T=39, L_q=80, L_ocr=256, L_obj=128, vocab=253, seed=1234)
        loss_of = lambda: model(batch)[1]                           # noqa: E731
        metric = "train samples/sec (PhonoSaL-large)"
        workload = f"PhonemeSaL T5-large, per-GPU batch {args.batch}, S=464, T=39"
    elif args.workload == "phonoprestu":
        _emit({"impl": "eager-gpu", "unavailable": "the oracle has no PhonemePreSTU module graph (that class is pinned "
                                                   "by fixtures recorded from the real reference, tests/golden)"})
        return
    else:
        cfg = ref_model.make_config()
        model = ref_model.PhonemeLaTr(cfg, 84, 187, 7).to(dev).train()
        batch = ref_model.synthetic_batch(args.batch, cfg, seed=1234)
        loss_of = lambda: ref_model.phoneme_latr_loss(model, batch, 2)       # noqa: E731
    optim = torch.optim.Adam(model.parameters(), lr=5e-5, betas=(0.9, 0.98), eps=1e-9)
    sched = torch.optim.lr_scheduler.LinearLR(optim, total_iters=2000)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    out = {}
    for name, amp in (("f32", False), ("bf16_autocast", True)):
        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                loss = loss_of()
            optim.zero_grad()
            loss.backward()
            optim.step()
            sched.step()
            return loss
        for _ in range(max(args.warmup, 1)):
            step()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            last = step()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / args.steps
        out[name] = {"samples_per_s": args.batch / (ms / 1e3), "ms_per_step": ms, "final_loss": float(last.item())}
    _emit({"impl": "eager-gpu", "metric": metric, "unit": "samples/s", "n_gpus": 1,
           "steps": args.steps, "warmup": max(args.warmup, 1), "data": "synthetic",
           "config": {"workload": workload + " — reference module graph, PyTorch eager"},
           "value": out["bf16_autocast"]["samples_per_s"], "dtype": "bf16 autocast", "variants": out})


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.cpu_batch, args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "train samples/sec (PhonoLaTr-base)", "value": res["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(B=args.cpu_batch) + " (CPU sample batch)"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch.distributed as dist
    import phoneme_vqa_b200 as pv
    from phoneme_vqa_b200 import models, ops, parallel, synthetic, train

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pv.load()

    B = args.batch
    cfg = synthetic.t5_config("base")
    torch.manual_seed(0)
    loss_fn, ignore_index = None, synthetic.PAD_ID
    shape = {}                            # sequence geometry for the algorithmic-bytes table (default: PhonoLaTr)
    metric = {"phonolatr": "train samples/sec (PhonoLaTr-base)", "latr": "train samples/sec (LaTr-base)",
              "phonoprestu": "train samples/sec (PhonoPreSTU-base)", "phonosal": "train samples/sec (PhonoSaL-large)"}[args.workload]
    if args.workload == "latr":          # BASELINE config 2 (not the headline metric; kept for completeness)
        cfg = synthetic.t5_config("base", num_decoder_layers=12)
        model = models.LaTr(cfg).to(dev)
        make_batch, ignore_index = synthetic.latr_batch, 0
        workload = f"LaTr T5-base (12+12 layers, 36096-way vocabulary head), per-GPU batch {B}, S=327, T=127"
    elif args.workload == "phonoprestu":  # BASELINE config 4: trainable ViT-B/16, question+OCR text in one sequence
        vit = None if args.image == 224 else dict(image_size=args.image)
        cfg = synthetic.t5_config("base", vit_config=vit)
        model = models.PhonemePreSTU(cfg, *synthetic.PHONEME_VOCAB).to(dev)
        make_batch = lambda B_, V, **kw: synthetic.phoneme_prestu_batch(B_, V, image=args.image, **kw)  # noqa: E731
        loss_fn = synthetic.phoneme_prestu_loss
        s_img = (args.image // 16) ** 2 + 1
        shape = dict(S_img=s_img, L_ocr=0, L_q=130)
        workload = (f"PhonemePreSTU T5-base, trainable ViT-B/16 at {args.image}px, per-GPU batch {B}, "
                    f"S={s_img + 130} ({s_img} ViT + 130 text), T=127, phoneme vocab 84/187/7")
    elif args.workload == "phonosal":     # BASELINE config 5: T5-large dims, 80 question + 256 OCR + 128 object tokens
        cfg = synthetic.t5_config("large")
        cfg.update({"ocr_hidden": 512, "obj_hidden": 2048, "new_token_embedding_size": cfg.vocab_size})
        model = models.PhonemeSaL(cfg, 253).to(dev)
        make_batch = synthetic.phoneme_sal_batch
        loss_fn = synthetic.phoneme_sal_loss(256, 80)
        shape = dict(S_img=0, L_ocr=256, L_q=80, T=39)
        workload = (f"PhonemeSaL T5-large (24 layers, 16 heads, 1-D + SCP bias in-kernel), per-GPU batch {B}, "
                    "S=464 (80 question + 256 OCR + 128 objects), T=39, 253-way phoneme vocabulary")
    else:
        model = models.PhonemeLaTr(cfg, *synthetic.PHONEME_VOCAB).to(dev)
        make_batch = synthetic.phoneme_latr_batch
        workload = WORKLOAD.format(B=B)
    model.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    model.train()
    ops.manual_seed(1234 + rank)
    reducer = parallel.GradReducer(model, bucket_mb=32.0, attn_bwd_waves=args.attn_bwd_waves)
    reducer.broadcast_parameters(0)
    trainer = train.TrainStep(model, reducer if world > 1 else None, lr=5e-5, betas=(0.9, 0.98), eps=1e-9,
                              warmup_iters=2000, ignore_index=ignore_index, use_graph=not args.no_graph,
                              loss_fn=loss_fn)

    n_distinct = 4
    host = [make_batch(B, cfg.vocab_size, seed=1234 + rank * 1000 + i, pin=True)
            for i in range(n_distinct)]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = synthetic.batch_bytes(host[0])

    def step(b):
        return trainer(b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host):
        barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = None
        for i in range(n_steps):
            if from_host:
                last = step(host[i % n_distinct]).item()   # pinned host -> device copies + loss read every step
            else:
                last = step(resident[i % n_distinct])
        e.record()
        barrier()
        ms = a.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for i in range(args.warmup):
        step(resident[i % n_distinct])
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0, replays0 = pv.launch_count(), trainer.replays
    ms, last = timed(args.steps, from_host=False)
    # kernels of libpvqa_sm100.so executed in the timed region: eager launches + (kernels per graph) x replays
    launches = (pv.launch_count() - launches0) + trainer.launches_per_replay * (trainer.replays - replays0)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, last_loss = timed(args.steps, from_host=True)

    # profiled pass: CUDA events around every C-ABI launch, same steps, eager (events cannot sit inside a graph)
    trainer.use_graph = False
    ops.KernelTimer.reset(True)
    timed(min(args.steps, 5), from_host=False)
    torch.cuda.synchronize()
    ksum = ops.KernelTimer.summary()
    ops.KernelTimer.reset(False)
    n_prof = min(args.steps, 5)

    if rank == 0:
        hbm, tf, src = _peaks()
        alg = kernel_algorithmic(B, cfg, **shape)
        traffic = _measured_traffic()
        kernels = {}
        for name, (n, total_ms) in ksum.items():
            avg_ms = total_ms / n
            entry = {"launches_per_step": n / n_prof, "avg_ms": avg_ms, "share_of_step": total_ms / n_prof / (ms / args.steps)}
            work = alg.get(name) or attention_flops(name, B, cfg)
            if work:
                kind, amount = work
                if kind == "hbm":
                    ach = amount / avg_ms / 1e6
                    entry.update({"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm})
                else:
                    ach = amount / avg_ms / 1e9
                    entry.update({"bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf})
            if name in traffic:
                entry["traffic"] = traffic[name]          # ncu dram bytes per launch (profiles/traffic.json)
                # the same launch rated by what ncu saw move to and from DRAM (cold L2) instead of by algorithmic bytes
                entry["dram_frac"] = traffic[name] / avg_ms / 1e6 / hbm
            kernels[name] = entry
        dom = max((k for k in kernels if "bound" in kernels[k]), key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches_per_step"],
                  default=None)
        roofline = None
        if dom:
            roofline = {k: kernels[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
            roofline.update({"kernel": dom, "traffic": traffic.get(dom), "peak_source": src,
                             "avg_launch_ms": kernels[dom]["avg_ms"],
                             "note": "achieved = algorithmic flops (or bytes) per launch / CUDA-event duration of that "
                                     "launch inside the step; traffic = ncu dram bytes per launch (profiles/)"})
        value = world * B * args.steps / (ms / 1e3)
        line = {
            "metric": metric, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload,
                       "global_batch": world * B,
                       "parallelism": f"dp{world}", "weights": "random-init", "attn_bwd_waves": reducer.attn_bwd_waves,
                       "eager_gpu_samples_s": _eager_gpu(args),
                       "launch": "eager" if args.no_graph else "one CUDA graph per step",
                       "l2": "no flush: per-step working set (0.9 GB weights+Adam state read, >10 GB activations) "
                             "exceeds the 126 MB L2; 4 distinct batches cycled"},
            "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "final_loss": float(last_loss),
        }
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference_run(args.cpu_batch, 1, 1)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _emit(line)
    if world > 1:
        # Orderly teardown: the captured graph (which holds NCCL kernels) goes first, then the communicator.  A watchdog
        # keeps the driver's scaling run from ever hanging at exit should the handshake stall (observed once in round 1
        # when the communicator was destroyed underneath a live graph).
        import gc
        import threading
        watchdog = threading.Timer(60.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        trainer.close()
        del trainer, reducer
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()


def _eager_gpu(args):
    """samples/s per GPU of the PyTorch-eager incumbent on a B200 for this workload, as measured by
    `bench.py --impl eager-gpu` and archived under profiles/ (None when this configuration was not measured)."""
    path = os.path.join(ROOT, "profiles", "eager_gpu.json")
    if not os.path.exists(path) or args.image != 224:
        return None
    with open(path) as f:
        w = json.load(f)["workloads"].get(args.workload)
    if not w:
        return None
    return {"f32": w["f32_samples_s"], "bf16_autocast": w["bf16_autocast_samples_s"], "n_gpus": 1, "source": w["source"]}


def _measured_traffic():
    """ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch at the bench shape, recorded under profiles/."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def attention_flops(name, B, cfg):
    """attn_{fwd,bwd}[Sq=..,Sk=..]: 4*B*H*Sq*Sk*D forward (2 GEMMs), 2.5x that backward (5 GEMMs);
    the causal decoder self-attention (Sq == Sk == T) only needs the lower triangle."""
    import re
    m = re.match(r"attn_(fwd|bwd)(?:_v[23])?\[Sq=(\d+),Sk=(\d+)\]", name)
    if not m:
        return None
    sq, sk = int(m.group(2)), int(m.group(3))
    fl = 4.0 * B * cfg.num_heads * sq * sk * cfg.d_kv
    if sq == sk and sq < 197:        # target-side self-attention is causal
        fl *= 0.5
    return ("tensor", fl if m.group(1) == "fwd" else 2.5 * fl)


def kernel_algorithmic(B, cfg, S_img=197, L_ocr=100, L_q=30, T=127):
    """algorithmic bytes / flops per launch of each C-ABI kernel.  Kernels
    that run at two shapes inside the step (encoder and target-decoder feed-forward) carry the launch-weighted mean,
    like the measured average they are divided by."""
    d, H, D = cfg.d_model, cfg.num_heads, cfg.d_kv
    S = S_img + L_ocr + L_q
    e_tab, e_act = 4, 2
    n_enc, n_dec, n_vit = B * S * d, B * T * d, B * S_img * d          # elements of one activation tensor
    L_enc, L_dec = cfg.num_layers, cfg.num_decoder_layers
    ff_enc, ff_dec = B * S * cfg.d_ff, B * T * 2048                     # nn.TransformerDecoderLayer default dim_feedforward

    def mix(enc, dec):
        return (L_enc * enc + L_dec * dec) / (L_enc + L_dec)

    N = B * T
    head_flops = 2.0 * N * d * d + 2.0 * N * (d // 3) * 278
    return {
        "embed_mm_fwd": ("hbm", B * ((7 * L_ocr + L_q) * (d * e_tab + 8) + S_img * d * e_act) + B * S * d * e_act),
        # gradient rows read once + indices + one fp32 reduction per (token, table) row (run-length merged in registers)
        "embed_mm_bwd": ("hbm", B * ((L_ocr + L_q) * d * e_act + (7 * L_ocr + L_q) * 8 + (7 * L_ocr + L_q) * d * 4)),
        "embed_tgt_fwd": ("hbm", B * T * (3 * 8 + d * 4 + d * 4 + d * 4)),
        "embed_tgt_bwd": ("hbm", B * T * (3 * 8 + d * 4 + 2 * d * 4)),
        "phoneme_head_ce_fwd": ("hbm", B * T * (d * e_act + 3 * 8 + 3 * 4)),
        "phoneme_head_ce_bwd": ("hbm", B * T * (d * e_act + 3 * 8 + 3 * 4 + 278 * e_act)),
        "phoneme_head_fused_fwd": ("tensor", head_flops),               # shared_lm_head GEMM + three head GEMMs
        # fp32 residual stream + bf16 update in, stream + bf16 normed row out  /  three gradients in, two out
        "add_dropout_rms_fwd": ("hbm", n_enc * (4 + 2 + 4 + 2)),
        "add_dropout_rms_bwd": ("hbm", n_enc * (2 + 4 + 4 + 4 + 2)),
        "add_dropout_ln_fwd": ("hbm", n_dec * (4 + 2 + 4 + 4 + 2)),
        "add_dropout_ln_bwd": ("hbm", n_dec * (4 + 2 + 4 + 4 + 2)),
        "add_ln_lp": ("hbm", n_vit * (2 + 2 + 2 + 2)),                  # frozen ViT: bf16 stream
        "relu_dropout_fwd": ("hbm", mix(ff_enc, ff_dec) * (2 + 2)),
        "relu_dropout_bwd": ("hbm", mix(ff_enc, ff_dec) * (2 + 2 + 2)),
        "cast_rows": ("hbm", mix(n_enc, n_dec) * (4 + 2)),
        "rms_norm_fwd": ("hbm", n_enc * (4 + 2)),
        "rms_norm_bwd": ("hbm", n_enc * (2 + 4 + 4 + 4)),
    }


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, HF warnings) print to stdout; the contract is ONE JSON line there.
    Point fd 1 at stderr for the run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager-gpu"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="phonolatr", choices=["phonolatr", "latr", "phonoprestu", "phonosal"],
                    help="phonolatr = the headline metric (BASELINE config 3's per-GPU shard); the others are the "
                         "sibling configs 2, 4 and 5")
    ap.add_argument("--image", type=int, default=224, choices=[224, 384], help="phonoprestu: ViT input size")
    ap.add_argument("--attn-bwd-waves", type=int, default=None,
                    help="N > 1: CTAs per SM of the attention backward (default: 1 up to 2 ranks, 4 beyond)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of one CUDA graph")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "eager-gpu":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl eager-gpu needs a CUDA device")
        run_eager_gpu_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        run_b200_arm(args)


if __name__ == "__main__":
    main()
