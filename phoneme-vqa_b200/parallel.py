"""Data-parallel training: one process per GPU, NCCL bucketed gradient all-reduce overlapped
with backward (the only collective on the path — SURVEY.md §8e; the reference has none).

`GradReducer` is a thin, dependency-free DDP: parameters are grouped in reverse registration
order (heads -> decoder -> encoder -> embeddings, the order gradients become ready) into flat
fp32 buckets of ~`bucket_mb`; a post-accumulate-grad hook copies each gradient into its bucket
slot and, when the bucket is full, launches an async all-reduce (NCCL runs it on its own
stream, so it overlaps the remaining backward); `finish()` waits, and leaves `p.grad` as views
into the averaged buckets (no copy back).  The trainable set can change between epochs (encoder
freeze toggle, core/executor/PhonemeLaTr_Executor.py:152-159): `rebuild()` re-buckets.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, average: bool = True):
        self.module = module
        self.bucket_bytes = int(bucket_mb * 1024 * 1024)
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._hooks = []
        self.rebuild()

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
        self._views[bi][pi].copy_(p.grad)
        p.grad = None
        self._filled.add((bi, pi))
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        flat = self._flat[bi]
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
