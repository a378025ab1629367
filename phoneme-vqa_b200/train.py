"""One training step of the reference's `_train_epoch` body (core/executor/PhonemeLaTr_Executor.py:161-198)
— forward, 3x cross-entropy, zero_grad, backward, (gradient all-reduce), Adam, LinearLR — optionally captured
into ONE CUDA graph so the ~3 000 kernel launches of a step cost one host call.

Graph-safety of the pieces:
  * dropout: host Philox offsets are baked into the captured launches; a device-resident step counter
    (pvqa_set_rng_step_counter) is added inside every dropout kernel and bumped by the graph itself, so each
    replay draws fresh masks;
  * learning rate: a device tensor (`capturable` fused Adam), refilled from the host LinearLR formula per step;
  * bf16 weight shadows: invalidated before capture so their refresh is part of the graph, and again after every
    replay (a replay updates the fp32 masters without bumping `_version`), so eager eval / generate between
    graph-trained steps never runs one optimizer step behind;
  * inputs: copied into static buffers before each replay; a batch whose shapes differ (short tail batch) runs eagerly;
  * trainable set: the reference toggles `encoder.requires_grad` per epoch (PhonemeLaTr_Executor.py:152-159) over an
    Adam built on ALL parameters (:262).  The `requires_grad` signature is recorded at capture; when it changes the
    graph and static buffers are dropped, the reducer re-buckets and the step is captured again.
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
        # reference: Adam(model.parameters(), lr, betas, eps=1e-9) + LinearLR(total_iters=warmup_step) stepped per
        # iteration (:262-266).  ALL parameters, frozen ones included (their grad is None, Adam skips them), so the
        # optimizer state_dict has the reference's single param group and survives the per-epoch freeze toggle.
        self.optim = torch.optim.Adam(list(model.parameters()), lr=self.lr_t, betas=betas, eps=eps, fused=True,
                                      capturable=True)
        self.iteration = 0
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.static_loss = None
        self.launches_per_replay = 0
        self.replays = 0
        self.captures = 0
        self._sig = None
        # the dropout step counter lives as long as the library (ops keeps it), not as long as this object
        self.rng_counter = ops.rng_step_counter(dev)

    def _signature(self):
        return tuple(p.requires_grad for p in self.model.parameters())

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
            self.reducer.step_overlapping_tail(self.optim)
        else:
            self.optim.step()
        return loss

    def eager(self, batch):
        loss = self._body(batch)
        self.iteration += 1
        self.lr_t.fill_(self._lr_now())
        return loss

    def capture(self, example_batch, warmup=3):
        """warm up eagerly on a side stream (allocator, cuBLAS workspaces, kernel attributes, Adam state), then capture.
        The warm-up must not train: it runs with lr = 0 (parameters unchanged), and the Adam moments / step counts it
        advanced, the iteration count and the dropout offset are put back afterwards, so the first replay is the first
        real step on this batch, as in the reference loop."""
        self.static = {k: v.clone() for k, v in example_batch.items()}
        it0, off0w = self.iteration, ops._Rng.offset
        saved = {}
        for p_, st in self.optim.state.items():
            saved[p_] = {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
        # ONE side stream for the warm-up and the capture: autograd pins every AccumulateGrad node to the stream it was
        # created on, and the gradient hooks of the reducer keep those nodes alive — created on another stream than the
        # capture runs on, every gradient would be handed over through a cross-stream sync inside the graph
        s = self._stream = getattr(self, "_stream", None) or torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            if self.reducer is not None and self.reducer.world > 1:
                self.reducer.rebuild()                  # re-register the hooks: fresh AccumulateGrad nodes, on `s`
            self.lr_t.zero_()
            for _ in range(warmup):
                self._body(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p_, st in self.optim.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if p_ in saved and k in saved[p_]:
                            v.copy_(saved[p_][k])
                        else:
                            v.zero_()
        self.iteration = it0
        ops._Rng.offset = off0w
        self.lr_t.fill_(self._lr_now())
        self._sig = self._signature()
        self.captures += 1
        SHADOWS.invalidate()
        off0 = ops._Rng.offset
        c0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=s):
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
        if self.graph is not None and self._signature() != self._sig:
            # the encoder freeze toggle changed the trainable set: the captured backward / Adam launches are stale
            self.graph, self.static, self.static_loss = None, None, None
            if self.reducer is not None:
                self.reducer.maybe_rebuild()
        if self.graph is not None and any(tuple(self.static[k].shape) != tuple(v.shape) for k, v in batch.items()):
            # short tail batch (DataLoader drop_last=False): never broadcast it into the static buffers
            loss = self.eager({k: v.to(self.device, non_blocking=True) for k, v in batch.items()})
            SHADOWS.invalidate()
            return loss
        if self.graph is None:
            self.capture({k: v.to(self.device) for k, v in batch.items()})
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        SHADOWS.invalidate()          # the replay moved the masters without bumping their _version
        self.replays += 1
        self.iteration += 1
        self.lr_t.fill_(self._lr_now())
        return self.static_loss

    def close(self):
        """Release the captured graph and its static buffers.  Call before dist.destroy_process_group(): the graph
        holds the captured NCCL kernels, and tearing the communicator down underneath it is what hung at exit."""
        self.graph, self.static, self.static_loss = None, None, None

    def check_indices(self):
        """Raise if any embedding kernel of this process saw an out-of-range token / coordinate / label index since
        the last check (the reference's nn.Embedding raises a device assert).  One .item() — call it per epoch or
        every N steps, never inside a capture."""
        ops.check_index_errors()
