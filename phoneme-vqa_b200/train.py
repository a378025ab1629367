"""One training step of the reference's `_train_epoch` body (core/executor/PhonemeLaTr_Executor.py:161-198)
— forward, 3x cross-entropy, zero_grad, backward, (gradient all-reduce), Adam, LinearLR — optionally captured
into ONE CUDA graph so the ~3 000 kernel launches of a step cost one host call.

Graph-safety of the pieces:
  * dropout: host Philox offsets are baked into the captured launches; a device-resident step counter
    (pvqa_set_rng_step_counter) is added inside every dropout kernel and bumped by the graph itself, so each
    replay draws fresh masks;
  * learning rate: a device tensor (`capturable` fused Adam), refilled from the host LinearLR formula per step;
  * bf16 weight shadows: invalidated before capture so their refresh is part of the graph;
  * inputs: copied into static buffers before each replay.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .modules import SHADOWS


class TrainStep:
    def __init__(self, model, reducer=None, lr=5e-5, betas=(0.9, 0.98), eps=1e-9, warmup_iters=2000,
                 start_factor=1.0 / 3.0, ignore_index=2, use_graph=True, loss_fn=None):
        """`loss_fn(model, batch) -> scalar loss` replaces the LaTr-family step body (PreSTU / SaL argument lists)."""
        self.model, self.reducer = model, reducer
        self.loss_fn = loss_fn
        self.base_lr, self.warmup_iters, self.start_factor = lr, warmup_iters, start_factor
        self.ignore_index = ignore_index
        dev = next(model.parameters()).device
        self.device = dev
        self.lr_t = torch.tensor(lr * start_factor, dtype=torch.float32, device=dev)
        params = [p for p in model.parameters() if p.requires_grad]
        # reference: Adam(lr, betas, eps=1e-9) + LinearLR(total_iters=warmup_step) stepped per iteration (:262-266)
        self.optim = torch.optim.Adam(params, lr=self.lr_t, betas=betas, eps=eps, fused=True, capturable=True)
        self.iteration = 0
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.static_loss = None
        self.launches_per_replay = 0
        self.replays = 0
        self.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.load().pvqa_set_rng_step_counter(self.rng_counter.data_ptr())

    def _lr_now(self):
        f = self.start_factor + (1.0 - self.start_factor) * min(self.iteration, self.warmup_iters) / self.warmup_iters
        return self.base_lr * f

    def _body(self, b):
        if self.loss_fn is not None:
            loss = self.loss_fn(self.model, b)
        else:
            labels = b["label_ids"]
            loss = self.model.forward_loss(b["pixel_values"], b["coordinates"], b["input_ids"], labels[:, :-1],
                                           b["src_attention_mask"], b["label_attention_mask"][:, :-1],
                                           b["ocr_attention_mask"], b["tokenized_ocr"], targets=labels[:, 1:],
                                           ignore_index=self.ignore_index)
        self.optim.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optim.step()
        return loss

    def eager(self, batch):
        loss = self._body(batch)
        self.iteration += 1
        self.lr_t.fill_(self._lr_now())
        return loss

    def capture(self, example_batch, warmup=3):
        """warm up eagerly on a side stream (allocator, cuBLAS workspaces, kernel attributes), then capture."""
        self.static = {k: v.clone() for k, v in example_batch.items()}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.eager(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        SHADOWS.invalidate()
        off0 = ops._Rng.offset
        c0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body(self.static)
            self.rng_counter.add_(ops._Rng.offset - off0)      # fresh dropout masks on every replay
        self.launches_per_replay = _lib.launch_count() - c0     # libpvqa kernels inside one replay
        torch.cuda.synchronize()
        return self

    def __call__(self, batch):
        """batch: dict of device (or pinned host) tensors.  Returns the loss tensor of this step."""
        if not self.use_graph:
            if next(iter(batch.values())).device != self.device:
                batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
            return self.eager(batch)
        if self.graph is None:
            self.capture({k: v.to(self.device) for k, v in batch.items()})
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.iteration += 1
        self.lr_t.fill_(self._lr_now())
        return self.static_loss
