"""Whole-step CUDA graph: forward + loss + backward + clip + AdamW captured once, replayed per batch.

The reference's step issues ~61 k ATen ops; ours is ~1.3 k launches, which at CIFAR batch sizes is
still launch-bound from Python.  Every C-ABI entry point is capture-safe (caller's stream, no allocation, no sync), so
the step is captured with torch.cuda.graph and replayed: the host's per-step work becomes one H2D copy of the
batch, one write of (lr, beta1, bias corrections) into pinned memory, and one graph launch."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .functional import cross_entropy


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, opt, example_x: torch.Tensor, example_y: torch.Tensor,
                 label_smoothing: float = 0.0, autocast_bf16: bool = True, warmup: int = 3,
                 before_backward: Optional[Callable[[], None]] = None, after_backward: Optional[Callable[[], None]] = None,
                 capture_error_mode: str = "global", between: Optional[Callable[[], None]] = None, mix: bool = False):
        """between: called eagerly between backward and clip + AdamW (the data-parallel gradient all-reduce).  The step is
        then two graphs -- [zero_grad, forward, loss, backward] and [clip, AdamW] -- with the collective outside both.
        mix: capture the two-target loss lam * CE(y) + (1 - lam) * CE(y_b) of the reference's CutMix / MixUp branch
        (H:1404-1408); y_b and lam are per-step inputs of __call__ (lam travels with the optimizer's per-step scalars)."""
        self.model, self.opt = model, opt
        self.x = example_x.clone()
        self.y = example_y.clone()
        self.mix = mix
        self.yb = example_y.clone() if mix else None
        self.ls, self.amp = label_smoothing, autocast_bf16
        self._bb, self._ab, self._between = before_backward, after_backward, between
        self.loss = torch.zeros((), device=self.x.device)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if mix:
                    self.opt.hyper[9] = 1.0
                self._step_body(eager=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # data-parallel runs capture the bucketed NCCL all-reduces too; the process group's watchdog thread issues CUDA
        # calls of its own, hence capture_error_mode="thread_local" there
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            if between is None:
                self._step_body(eager=False)
            else:
                self._fwd_bwd()
        self.graph2 = None
        if between is not None:
            torch.cuda.synchronize()
            self.graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph2, pool=self.graph.pool(), capture_error_mode=capture_error_mode):
                self._update()
        torch.cuda.synchronize()

    def _fwd_bwd(self):
        self.opt.zero_grad()
        if self._bb:
            self._bb()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            logits = self.model(self.x)
        if self.mix:
            loss = cross_entropy(logits, self.y, label_smoothing=self.ls, target_b=self.yb, lam=self.opt.hyper[9])
        else:
            loss = cross_entropy(logits, self.y, label_smoothing=self.ls)
        loss.backward()
        if self._ab:
            self._ab()
        self.loss.copy_(loss.detach())

    def _update(self):
        self.opt.clip()
        self.opt.step()

    def _step_body(self, eager: bool):
        self._fwd_bwd()
        if eager and self._between:
            self._between()
        self._update()

    def prefetch(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Start the host -> device copy of the NEXT step's batch (pinned host tensors) on a copy stream, into a staging pair; the
        next ``__call__()`` without arguments consumes it.  Called right after launching step i, the copy of batch i + 1 runs under
        step i instead of in front of step i + 1 (58 MB per step at the headline batch: ~2 % of the step)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._xs, self._ys = torch.empty_like(self.x), torch.empty_like(self.y)
            self._stage_ready, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.x.device))
            self._pending = False
        cs = self._copy_stream
        cs.wait_event(self._stage_free)          # the step that consumed the previous staging pair has read it
        with torch.cuda.stream(cs):
            self._xs.copy_(x, non_blocking=True)
            self._ys.copy_(y, non_blocking=True)
            self._stage_ready.record(cs)
        self._pending = True

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None,
                 y_b: Optional[torch.Tensor] = None, lam: float = 1.0) -> torch.Tensor:
        """Run one training step on (x, y) (host or device tensors; None = the prefetched batch if prefetch() was called, else the
        resident batch).  With mix=True, (y_b, lam) select the two-target loss of this step (y_b None: plain CE, lam = 1)."""
        if x is None and getattr(self, "_pending", False):
            main = torch.cuda.current_stream(self.x.device)
            main.wait_event(self._stage_ready)
            self.x.copy_(self._xs, non_blocking=True)      # device -> device into the graph's static input
            self.y.copy_(self._ys, non_blocking=True)
            self._stage_free.record(main)
            self._pending = False
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if self.mix:
            if y_b is not None:
                self.yb.copy_(y_b, non_blocking=True)
            else:
                lam = 1.0
        elif y_b is not None:
            raise RuntimeError("GraphedTrainStep: build with mix=True to use a second target")
        self.opt.push_hyper(lam)         # lr / beta1 / bias corrections of THIS step -> device buffer (pinned ring + H2D copy
        self.graph.replay()             #   in stream order BEFORE the replay; the graph itself holds no H2D node)
        if self.graph2 is not None:
            self._between()
            self.graph2.replay()
        return self.loss
