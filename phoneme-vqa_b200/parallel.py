"""Data-parallel training: one process per GPU, NCCL bucketed gradient all-reduce overlapped
with backward (the only collective on the path — SURVEY.md §8e; the reference has none).

`GradReducer` is a thin, dependency-free DDP: parameters are grouped in reverse registration
order (heads -> decoder -> encoder -> embeddings, the order gradients become ready) into flat
fp32 buckets of ~`bucket_mb`; when the last gradient of a bucket is ready an async all-reduce
is launched (NCCL runs it on its own stream, so it overlaps the remaining backward); `finish()`
waits, and leaves `p.grad` as views into the averaged buckets (no copy back).
Per-step passes over the 600 MB of gradients that round 1 paid and this version does not:
  * the weight-gradient GEMMs of the linears write straight INTO their bucket slot
    (`grad_slot()`, used by modules._LinearLP) — no gradient -> bucket `copy_` for ~95 % of the bytes;
  * the average is taken by the collective itself (`ReduceOp.AVG` on NCCL) — no `div_` pass.  The trainable set can change between epochs (encoder
freeze toggle, core/executor/PhonemeLaTr_Executor.py:152-159): `rebuild()` re-buckets.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_ACTIVE = None          # the reducer whose bucket slots gradient producers may write into directly


def grad_slot(param):
    """The bucket view a producer may write `param`'s gradient into (fp32, param's shape), or None.  A producer that
    uses it returns the view itself as the gradient: autograd then adopts it as `param.grad` without a copy."""
    r = _ACTIVE
    if r is None or r.world == 1:
        return None
    loc = r._slot.get(id(param))
    if loc is None:
        return None
    return r._views[loc[0]][loc[1]]


class GradReducer:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, average: bool = True):
        self.module = module
        self.bucket_bytes = int(bucket_mb * 1024 * 1024)
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: divide, then sum
        self._native_avg = bool(average and dist.is_initialized() and dist.get_backend(process_group) == "nccl")
        self._hooks = []
        self.rebuild()
        global _ACTIVE
        _ACTIVE = self

    # -- setup -------------------------------------------------------------------
    def broadcast_parameters(self, src: int = 0):
        if self.world == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src=src, group=self.pg)

    def rebuild(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        seen, params = set(), []
        for p in self.module.parameters():
            if p.requires_grad and id(p) not in seen:      # tied weights appear once
                seen.add(id(p))
                params.append(p)
        params.reverse()
        self.trainable_signature = tuple(id(p) for p in params)
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in params:
            nb = p.numel() * 4
            if cur and cur_bytes + nb > self.bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self.buckets.append(cur)
        self._flat, self._views, self._slot = [], [], {}
        for bi, bucket in enumerate(self.buckets):
            dev = bucket[0].device
            flat = torch.zeros(sum(p.numel() for p in bucket), dtype=torch.float32, device=dev)
            views, off = [], 0
            for pi, p in enumerate(bucket):
                views.append(flat[off:off + p.numel()].view_as(p))
                self._slot[id(p)] = (bi, pi)
                off += p.numel()
            self._flat.append(flat)
            self._views.append(views)
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._filled = set()
        if self.world > 1:
            for p in params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def maybe_rebuild(self):
        """call at epoch boundaries: re-bucket if requires_grad flags changed."""
        seen, sig = set(), []
        for p in self.module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                sig.append(id(p))
        sig.reverse()
        if tuple(sig) != self.trainable_signature:
            self.rebuild()
            return True
        return False

    # -- per-step ------------------------------------------------------------------
    def _on_grad(self, p):
        bi, pi = self._slot[id(p)]
        view = self._views[bi][pi]
        if p.grad.data_ptr() != view.data_ptr():        # (a producer that wrote into the slot already is a no-op here)
            view.copy_(p.grad)
        p.grad = None
        self._filled.add((bi, pi))
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        flat = self._flat[bi]
        if self._native_avg:
            self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            return
        if self.average:
            flat.div_(self.world)
        self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def finish(self):
        """wait for all buckets and expose the averaged gradients as p.grad (bucket views)."""
        if self.world == 1:
            return
        for bi, bucket in enumerate(self.buckets):
            if self._pending[bi] != 0:
                # parameters that received no gradient this step contribute zeros
                for pi in range(len(bucket)):
                    if (bi, pi) not in self._filled:
                        self._views[bi][pi].zero_()
                self._launch(bi)
        for w in self._works:
            if w is not None:
                w.wait()
        for bi, bucket in enumerate(self.buckets):
            for pi, p in enumerate(bucket):
                p.grad = self._views[bi][pi]
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._filled = set()
