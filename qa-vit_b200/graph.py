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

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None,
                 y_b: Optional[torch.Tensor] = None, lam: float = 1.0) -> torch.Tensor:
        """Run one training step on (x, y) (host or device tensors; None = reuse the resident batch).  With mix=True,
        (y_b, lam) select the two-target loss of this step (y_b None: plain CE, lam = 1)."""
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
