"""Data-parallel training: one process per GPU, NCCL gradient all-reduce overlapped with backward
(the only collective on the path — SURVEY.md §8e; the reference has none).

`GradReducer` is a thin, dependency-free DDP: flat fp32 buckets of ~`bucket_mb` with 256-byte aligned
slots; when the last gradient of a bucket is ready an async all-reduce is launched (NCCL runs it on its own stream, so
it overlaps the remaining backward); `finish()` / `step_overlapping_tail()` wait and leave `p.grad` as views into the
averaged buckets (no copy back).
  * the weight-gradient GEMMs of the linears write straight INTO their bucket slot (`grad_slot()`, used by
    modules._LinearLP) — no gradient -> bucket `copy_` for ~95 % of the bytes;
  * the average is taken by the collective itself (`ReduceOp.AVG` on NCCL) — no `div_` pass;
  * buckets start in reverse registration order and are rebuilt, after the first backward, in the order the gradients
    really became ready (rank 0's order, broadcast); parameters the loss never reaches are learned in the next step;
  * the last `tail_mb` in that order (the embedding tables) go over a second communicator, and the optimizer updates
    everything else while they are on the wire.
The trainable set can change between epochs (encoder freeze toggle, core/executor/PhonemeLaTr_Executor.py:152-159):
`maybe_rebuild()` re-buckets.  Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
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
    if loc is None or loc in r._handed_out:
        return None          # a second producer of the same parameter in one step must not overwrite the first
    r._handed_out.add(loc)
    return r._views[loc[0]][loc[1]]


class GradReducer:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, average: bool = True,
                 tail_mb: float = 160.0, attn_bwd_waves: int | None = None):
        """tail_mb: the last `tail_mb` of gradients in launch order (the embedding tables: the backward produces them
        last, nothing is left to hide their all-reduce behind) are the "tail"; step_overlapping_tail() updates every
        other parameter while the tail is still on the wire.
        attn_bwd_waves: CTAs per SM the persistent attention backward is cut into while collectives share the GPU
        (pvqa_set_attn_bwd_waves).  Default: 1 up to 2 ranks, 4 beyond (profiles/r02_ddp_timeline_n8.txt)."""
        self.module = module
        self.bucket_bytes = int(bucket_mb * 1024 * 1024)
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        nccl = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: divide, then sum
        self._native_avg = bool(average and nccl)
        self.tail_bytes = int(tail_mb * 1024 * 1024)
        # the tail travels on its own communicator (own NCCL stream): waiting for the head buckets must not mean
        # waiting for everything that was queued behind them on one in-order stream
        self.pg_tail = dist.new_group(backend="nccl") if (nccl and self.world > 1) else process_group
        self.attn_bwd_waves = 1
        if nccl and self.world > 1:
            self.attn_bwd_waves = int(attn_bwd_waves) if attn_bwd_waves is not None else (4 if self.world > 2 else 1)
            from . import _lib
            _lib.check(_lib.load().pvqa_set_attn_bwd_waves(self.attn_bwd_waves), "pvqa_set_attn_bwd_waves")
        self._hooks = []
        self._unused = None          # (bucket, slot) pairs that received no gradient in the first step
        self.rebuild()
        global _ACTIVE
        _ACTIVE = self

    def close(self):
        """undo the process-wide launch setting (tests build several reducers in one process)"""
        if self.attn_bwd_waves != 1:
            from . import _lib
            _lib.load().pvqa_set_attn_bwd_waves(1)
            self.attn_bwd_waves = 1

    # -- setup -------------------------------------------------------------------
    def broadcast_parameters(self, src: int = 0):
        if self.world == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src=src, group=self.pg)

    def rebuild(self, order=None):
        """order: positions into the canonical (reverse-registration) parameter list, the sequence in which the
        gradients became ready in a real backward.  None = canonical order, and the first step learns the real one."""
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
        self._canonical = list(params)
        self._arrival = None if order is not None else []   # [] = learning
        if order is not None:
            params = [params[i] for i in order]
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
            # every slot starts on a 256-byte boundary: the weight-gradient GEMMs write straight into them, and a
            # 4-byte-aligned output sends cuBLAS to its slow `align2` kernels (measured: 4 GEMMs, 70 -> 360 us/step)
            pad = lambda n: (n + 63) // 64 * 64                               # noqa: E731
            flat = torch.zeros(sum(pad(p.numel()) for p in bucket), dtype=torch.float32, device=dev)
            views, off = [], 0
            for pi, p in enumerate(bucket):
                views.append(flat[off:off + p.numel()].view_as(p))
                self._slot[id(p)] = (bi, pi)
                off += pad(p.numel())
            self._flat.append(flat)
            self._views.append(views)
        # launch order == bucket order; the last tail_bytes of it have no backward left to hide behind
        self._tail = [False] * len(self.buckets)
        acc = 0
        for bi in range(len(self.buckets) - 1, -1, -1):
            acc += self._flat[bi].numel() * 4
            self._tail[bi] = True
            if acc >= self.tail_bytes:
                break
        self._unused = None
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._filled = set()
        self._handed_out = set()
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
        if self._unused and (bi, pi) in self._unused:
            raise RuntimeError("GradReducer: a parameter that received no gradient in the first step received one now; "
                               "its bucket may already be on the wire.  Call rebuild() when the used set changes.")
        if self._arrival is not None:
            self._arrival.append(id(p))
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
        pg = self.pg_tail if self._tail[bi] else self.pg
        if self._native_avg:
            self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=pg, async_op=True)
            return
        if self.average:
            flat.div_(self.world)
        self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=pg, async_op=True)

    def _flush_unlaunched(self):
        """Buckets still waiting for a gradient at the end of the backward hold parameters the loss does not reach
        (e.g. layer 0's relative_attention_bias under an external position bias).  The first step finds them here
        and launches those buckets late — behind the tail on the in-order NCCL stream; from then on they are known
        (`_unused`), contribute zeros, and their buckets go out as soon as the live gradients are in."""
        late = [bi for bi in range(len(self.buckets)) if self._pending[bi] != 0]
        if self._unused is None:
            self._unused = {(bi, pi) for bi in late for pi in range(len(self.buckets[bi])) if (bi, pi) not in self._filled}
        for bi in late:
            for pi in range(len(self.buckets[bi])):
                if (bi, pi) not in self._filled:
                    self._views[bi][pi].zero_()
            self._launch(bi)
            self._pending[bi] = 0

    def _expose(self, tail):
        for bi, bucket in enumerate(self.buckets):
            if self._tail[bi] == tail:
                if self._works[bi] is not None:
                    self._works[bi].wait()
                for pi, p in enumerate(bucket):
                    p.grad = self._views[bi][pi]

    def _adopt_arrival_order(self):
        """After the first real backward: re-bucket in the order the gradients actually became ready (registration
        order is a poor guess: the layout tables and the target-side embeddings are registered next to the heads but
        their gradients are the last thing the backward produces, and a bucket goes out when its LAST gradient is in).
        Rank 0's order is used everywhere, so every rank builds identical buckets."""
        pos = {id(p): i for i, p in enumerate(self._canonical)}
        order = [pos[i] for i in self._arrival]
        got = set(order)
        order += [i for i in range(len(self._canonical)) if i not in got]      # never reached by the loss: at the end
        t = torch.tensor(order, dtype=torch.int64, device=self._flat[0].device)
        dist.broadcast(t, src=0, group=self.pg)
        self.rebuild(order=[int(i) for i in t.tolist()])

    def _reset(self):
        self._handed_out = set()
        if self._arrival:                                  # the learning step just ended
            if not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
                self._adopt_arrival_order()                # (never inside a capture: it broadcasts and reallocates)
                return
            self._arrival = None
        unused = self._unused or ()
        self._pending = [len(b) - sum(1 for pi in range(len(b)) if (bi, pi) in unused) for bi, b in enumerate(self.buckets)]
        self._works = [None] * len(self.buckets)
        self._filled = set()

    def finish(self):
        """wait for all buckets and expose the averaged gradients as p.grad (bucket views)."""
        if self.world == 1:
            return
        self._flush_unlaunched()
        self._expose(False)
        self._expose(True)
        self._reset()

    def step_overlapping_tail(self, optimizer):
        """finish() + optimizer.step(), with the update of everything outside the tail buckets running WHILE the tail
        all-reduce (the embedding tables: their gradient is the last thing the backward produces, so nothing else is
        left to hide it behind) is still on the wire.  Adam-family optimizers skip parameters whose grad is None and
        keep a step count per parameter, so two step() calls over disjoint sets are exactly one step() over the union."""
        if self.world == 1:
            optimizer.step()
            return
        self._flush_unlaunched()
        self._expose(False)
        optimizer.step()                                # tail parameters: grad is None here
        head = [p for bi, bucket in enumerate(self.buckets) if not self._tail[bi] for p in bucket]
        for p in head:
            p.grad = None
        self._expose(True)
        optimizer.step()
        for bi, bucket in enumerate(self.buckets):      # leave every averaged gradient visible, as finish() does
            if not self._tail[bi]:
                for pi, p in enumerate(bucket):
                    p.grad = self._views[bi][pi]
        self._reset()
