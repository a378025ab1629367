"""GradReducer (bucketed gradient all-reduce overlapped with backward) on 2 gloo ranks, CPU."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _make_model():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                            torch.nn.Linear(32, 4))
    m.frozen = torch.nn.Linear(4, 4)
    for p in m.frozen.parameters():
        p.requires_grad = False
    return m


def _worker(rank, world, store_path, q):
    # file rendezvous: no port to race for when several test sessions share the machine
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    dist.init_process_group("gloo", init_method="file://" + store_path, rank=rank, world_size=world)
    import phoneme_vqa_b200.parallel as par
    model = _make_model()
    if rank == 1:                                    # ranks start different: broadcast must fix it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    red = par.GradReducer(model, bucket_mb=0.002)    # tiny buckets -> several of them
    red.broadcast_parameters(0)
    assert len(red.buckets) >= 3
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    loss = torch.nn.functional.mse_loss(model(xs), ys)
    loss.backward()
    red.finish()
    grads = [p.grad.numpy().copy() for p in model.parameters() if p.requires_grad]   # numpy: pickled by value
    # toggle the trainable set (encoder freeze, PhonemeLaTr_Executor.py:152-159) and make sure re-bucketing works
    for p in model[0].parameters():
        p.requires_grad = False
    assert red.maybe_rebuild()
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(xs), ys)
    loss.backward()
    red.finish()
    grads2 = [p.grad.numpy().copy() for p in model.parameters() if p.requires_grad]
    if rank == 0:
        q.put((grads, grads2))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, 2, os.path.join(tmp, "store"), q)) for r in range(2)]
        for p in procs:
            p.start()
        grads, grads2 = q.get(timeout=300)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    model = _make_model()
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    torch.nn.functional.mse_loss(model(x), y).backward()          # equal shard sizes: mean of means == global mean
    ref = [p.grad for p in model.parameters() if p.requires_grad]
    assert len(ref) == len(grads)
    for a, b in zip(grads, ref):
        torch.testing.assert_close(torch.from_numpy(a), b, rtol=1e-5, atol=1e-6)
    ref2 = [p.grad for p in list(model.parameters())[2:] if p.requires_grad]
    for a, b in zip(grads2, ref2):
        torch.testing.assert_close(torch.from_numpy(a), b, rtol=1e-5, atol=1e-6)


def _worker_tail(rank, world, store_path, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    dist.init_process_group("gloo", init_method="file://" + store_path, rank=rank, world_size=world)
    import phoneme_vqa_b200.parallel as par
    model = _make_model()
    red = par.GradReducer(model, bucket_mb=0.002, tail_mb=0.003)     # several buckets, the last two are "tail"
    red.broadcast_parameters(0)
    assert any(red._tail) and not all(red._tail)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.98), eps=1e-9)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(xs), ys).backward()
        red.step_overlapping_tail(opt)
        assert all(p.grad is not None for p in model.parameters() if p.requires_grad)
    if rank == 0:
        q.put(([p.detach().numpy().copy() for p in model.parameters()],      # numpy: pickled by value
               [int(opt.state[p]["step"]) for p in model.parameters() if p.requires_grad]))
    dist.barrier()
    dist.destroy_process_group()


def test_split_optimizer_step_around_the_tail_buckets_is_one_adam_step():
    """step_overlapping_tail = two optimizer.step() calls over disjoint parameter sets (everything but the tail
    buckets first, the tail once its all-reduce has landed): must equal single-process Adam on the global batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker_tail, args=(r, 2, os.path.join(tmp, "store"), q)) for r in range(2)]
        for p in procs:
            p.start()
        params, steps = q.get(timeout=300)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    assert steps == [3] * len(steps)                                  # every parameter stepped exactly once per iteration
    model = _make_model()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.98), eps=1e-9)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(x), y).backward()
        opt.step()
    for a, b in zip(params, model.parameters()):
        torch.testing.assert_close(torch.from_numpy(a), b.detach(), rtol=1e-4, atol=1e-6)


class _Scrambled(torch.nn.Module):
    """registration order deliberately unlike the order gradients become ready: the LAST layer of the forward is
    registered first (like the layout tables of PhonemeLaTr, registered next to the heads, produced last by backward)"""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.first_in_forward = torch.nn.Linear(16, 32)      # its gradient arrives LAST
        self.head = torch.nn.Linear(32, 4)                    # its gradient arrives FIRST
        self.middle = torch.nn.Linear(32, 32)
        self.never_used = torch.nn.Linear(3, 3)               # the loss does not reach it

    def forward(self, x):
        return self.head(torch.relu(self.middle(torch.relu(self.first_in_forward(x)))))


def _worker_order(rank, world, store_path, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    dist.init_process_group("gloo", init_method="file://" + store_path, rank=rank, world_size=world)
    import phoneme_vqa_b200.parallel as par
    model = _Scrambled()
    red = par.GradReducer(model, bucket_mb=0.003, tail_mb=0.003)
    red.broadcast_parameters(0)
    names = {id(p): n for n, p in model.named_parameters()}
    before = [names[id(p)] for b in red.buckets for p in b]
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    grads = None
    for step in range(3):
        model.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(xs), ys).backward()
        red.finish()
        if step == 0:
            after = [names[id(p)] for b in red.buckets for p in b]
        grads = {n: p.grad.numpy().copy() for n, p in model.named_parameters()}
    unused = sorted(names[id(red.buckets[bi][pi])] for bi, pi in (red._unused or ()))
    tail = [names[id(p)] for bi, b in enumerate(red.buckets) if red._tail[bi] for p in b]
    if rank == 0:
        q.put((before, after, unused, tail, grads))
    dist.barrier()
    dist.destroy_process_group()


def test_buckets_follow_gradient_arrival_order_after_the_first_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker_order, args=(r, 2, os.path.join(tmp, "store"), q)) for r in range(2)]
        for p in procs:
            p.start()
        before, after, unused, tail, grads = q.get(timeout=300)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    used = [n for n in after if not n.startswith("never_used")]
    # arrival order of a backward: head, then middle, then the first layer of the forward; unreachable ones last
    assert [n.split(".")[0] for n in used] == ["head", "head", "middle", "middle", "first_in_forward", "first_in_forward"]
    assert [n.split(".")[0] for n in after[-2:]] == ["never_used", "never_used"]
    assert before != after                                   # registration order had the first layer in the middle
    assert unused == ["never_used.bias", "never_used.weight"]
    # the tail is the END of the arrival order (at least tail_mb of it): the last-arriving gradients, never the heads
    assert tail == after[-len(tail):] and "first_in_forward.weight" in tail and "head.weight" not in tail
    # and the numbers are still the global-batch gradients; unreachable parameters get zeros (DDP semantics)
    model = _Scrambled()
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    torch.nn.functional.mse_loss(model(x), y).backward()
    for n, p in model.named_parameters():
        ref = p.grad if p.grad is not None else torch.zeros_like(p)
        torch.testing.assert_close(torch.from_numpy(grads[n]), ref, rtol=1e-5, atol=1e-6)
