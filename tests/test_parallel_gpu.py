"""Data-parallel gradient exchange over NCCL: two ranks, each on its own GPU, run the
CUDA-graph-captured TrainStep with parallel.GradReducer on their shard; the averaged gradients every rank ends up
with must equal the mean of the gradients ONE process computes on each shard in turn (data-parallel semantics).  Skipped on boxes with fewer than two GPUs (the
gloo version of the same check runs on CPU in test_parallel_cpu.py)."""
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
VOCAB = (21, 33, 7)


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        for attr in ("dropout", "p"):
            if isinstance(getattr(m, attr, None), float):
                setattr(m, attr, 0.0)


def _model_and_batch(dev):
    import phoneme_vqa_b200.models as M
    from oracle import ref_model
    cfg = ref_model.tiny_config()
    oracle = ref_model.PhonemeLaTr(cfg, *VOCAB)
    sd = ref_model.deterministic_state_dict(oracle)
    model = M.PhonemeLaTr(cfg, *VOCAB)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    model.train(); _no_dropout(model)
    model.set_compute_dtype(torch.bfloat16)
    batch = ref_model.synthetic_batch(4, cfg, T=17, L_ocr=20, L_q=8, V_sub=VOCAB, seed=11, image=32)
    return model, batch


def _worker(rank, world, store_path, out_path):
    import torch.distributed as dist
    from phoneme_vqa_b200 import parallel, train
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", store=dist.FileStore(store_path, world), rank=rank, world_size=world)
    model, batch = _model_and_batch(dev)
    shard = {k: v[rank * 2:(rank + 1) * 2].to(dev) for k, v in batch.items()}
    reducer = parallel.GradReducer(model, bucket_mb=0.25)           # several buckets even on the tiny model
    step = train.TrainStep(model, reducer, lr=0.0, warmup_iters=1, ignore_index=2, use_graph=True)
    step(shard)                                                     # capture + first replay
    step(shard)                                                     # a pure replay
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        torch.save(grads, out_path)
    step.close()                                                    # captured NCCL work must go before the communicator
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_graph_captured_nccl_gradients_equal_the_global_batch():
    import torch.multiprocessing as mp
    tmp = tempfile.mkdtemp()
    out = os.path.join(tmp, "grads.pt")
    mp.spawn(_worker, args=(2, os.path.join(tmp, "store"), out), nprocs=2, join=True)
    got = torch.load(out)
    # data-parallel semantics (what torch DDP computes): the mean over ranks of each rank's own mean-token loss
    # gradient.  (Not the gradient of the global batch's token mean: the shards hold different numbers of targets.)
    model, batch = _model_and_batch(torch.device("cuda", 0))
    ref = None
    for r in range(2):
        b = {k: v[r * 2:(r + 1) * 2].to("cuda:0") for k, v in batch.items()}
        labels = b["label_ids"]
        model.zero_grad(set_to_none=True)
        loss = model.forward_loss(b["pixel_values"], b["coordinates"], b["input_ids"], labels[:, :-1], b["src_attention_mask"],
                                  b["label_attention_mask"][:, :-1], b["ocr_attention_mask"], b["tokenized_ocr"],
                                  targets=labels[:, 1:], ignore_index=2)
        loss.backward()
        g = {k: p.grad.detach().float().cpu() / 2 for k, p in model.named_parameters() if p.grad is not None}
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    assert set(ref) == set(got)
    worst = max(((k, float((got[k] - ref[k]).norm() / (ref[k].norm() + 1e-12))) for k in ref), key=lambda t: t[1])
    print(f"[2-rank NCCL, graph-captured] worst relative gradient difference to the mean of the per-shard gradients: {worst}")
    assert worst[1] <= 2e-3, worst            # same kernels on the same shards: only atomics order and the fp32 AVG differ
