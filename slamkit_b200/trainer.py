"""Data-parallel gradient synchronisation and the train-step driver for hot path (ii).

Replaces what HF Trainer + accelerate's DDP wrapper do around `training_step` (HF:trainer.py:1867-2014;
config/training_args/default.yaml:18 `ddp_find_unused_parameters: false`): one process per GPU, each rank computes
gradients normalised by the GLOBAL number of label tokens, then a SUM all-reduce over NVLink (NCCL via
torch.distributed -- the only collective on the path) and the optimiser step.  The flat bf16 gradient buffer is reduced
in a few large buckets; bucket k's all-reduce is enqueued on a side stream as soon as the backward pass has finished
the layers it covers (CUDA events recorded inside `sk_lm_forward_backward`), so communication overlaps the rest of the
backward pass.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib as L
from .lm import B200AdamW, B200UnitLM, cosine_with_min_lr


class GradSync:
    def __init__(self, model: B200UnitLM, layers_per_bucket: int = 4, overlap: bool = True):
        import torch.distributed as dist
        self.dist = dist
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        nl = model.config.n_layers
        t = model.tensors
        self.layer_start = [t[f"layers.{l}.ln1"][0] for l in range(nl)] + [t["final_norm"][0]]
        self.buckets: List[Tuple[int, int, int]] = []   # (event index to wait for, start elem, end elem)
        hi = nl
        while hi > 0:
            lo = max(0, hi - layers_per_bucket)
            self.buckets.append((lo, self.layer_start[lo], self.layer_start[hi]))
            hi = lo
        self.tail = (self.layer_start[nl], model.n_params)   # final_norm + (tied) embedding: complete at the very end
        self.events: List[torch.cuda.Event] = []
        if self.overlap:
            self.comm = torch.cuda.Stream(device=model.device)
            self.events = [torch.cuda.Event() for _ in range(nl + 1)]
            for e in self.events:
                e.record()                      # forces creation of the underlying cudaEvent_t
            arr = (C.c_void_p * (nl + 1))(*[C.c_void_p(e.cuda_event) for e in self.events])
            L.check(model.lib.sk_lm_set_backward_events(model._h, arr, nl + 1))

    def reduce(self) -> None:
        """Call right after `forward_backward` of the LAST micro-batch of the accumulation window has been enqueued."""
        if self.world == 1:
            return
        g = self.model.grads
        if not self.overlap:
            self.dist.all_reduce(g)
            return
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.comm):
            for ev_idx, lo, hi in self.buckets:
                self.comm.wait_event(self.events[ev_idx])
                self.dist.all_reduce(g[lo:hi])
            self.comm.wait_stream(cur)
            self.dist.all_reduce(g[self.tail[0]:self.tail[1]])
        cur.wait_stream(self.comm)


class B200Trainer:
    """Minimal equivalent of `SLAMTrainer.train()`'s inner loop for the unit-LM recipe: gradient accumulation with HF
    `num_items_in_batch` semantics, clip 0.5, AdamW, `cosine_with_min_lr` schedule, token counting
    (slamkit/trainer/slam_trainer.py:59-71) done on-device and read only when asked."""

    def __init__(self, model: B200UnitLM, lr: float = 1e-3, min_lr: float = 5e-5, warmup_steps: int = 100,
                 total_steps: int = 17625, max_grad_norm: float = 0.5, weight_decay: float = 0.0,
                 grad_accum: int = 1, overlap_comm: bool = True):
        self.model = model
        self.opt = B200AdamW(model, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.sync = GradSync(model, overlap=overlap_comm)
        self.lr, self.min_lr, self.warmup, self.total = lr, min_lr, warmup_steps, total_steps
        self.grad_accum = grad_accum
        self.step_idx = 0
        self.tokens_seen = torch.zeros((), device=model.device, dtype=torch.float64)

    def train_step(self, micro_batches) -> torch.Tensor:
        """micro_batches: list of dicts with input_ids/labels(/position_ids). Returns the device loss of the window."""
        assert len(micro_batches) == self.grad_accum
        n_items = sum(float((mb["labels"] != -100).sum()) for mb in micro_batches)
        if self.sync.world > 1:
            t = torch.tensor([n_items], device=self.model.device, dtype=torch.float64)
            self.sync.dist.all_reduce(t)
            n_items = float(t.item())
        loss = torch.zeros((), device=self.model.device)
        for i, mb in enumerate(micro_batches):
            out = self.model.forward_backward(mb["input_ids"], mb["labels"], mb.get("position_ids"),
                                              num_items_in_batch=n_items, accumulate=i > 0)
            loss = loss + out.stats[0]
            self.tokens_seen += out.stats[1].double()
        self.sync.reduce()
        lr = cosine_with_min_lr(self.step_idx, base_lr=self.lr, min_lr=self.min_lr, warmup_steps=self.warmup,
                                total_steps=self.total)
        self.opt.step(lr=lr)
        self.step_idx += 1
        return loss
