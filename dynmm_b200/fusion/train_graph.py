"""One CUDA graph for a whole training step (forward + backward + gradient exchange + optimizer).

A training step of the gated RGB-D network is >3000 small launches (every convolution is a forward, a
data-gradient and a weight-gradient kernel, plus BatchNorm / ReLU / optimizer element-wise work); issued
one by one from Python the step is host-bound (~60 ms at batch 8) whatever the kernels cost.  All of our
entry points are stream-ordered, allocation-free and sync-free, so the whole step captures into one graph
and replays at device speed.  The soft / hard gate needs no host decision in training (every sample's
depth features are required for the gate gradient), so one graph serves all gate
outcomes.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

Tensor = torch.Tensor


class GraphedTrainStep:
    """``step(rgb, depth, target) -> loss`` replaying one captured graph.

    ``loss_fn(model_output, target) -> scalar`` where ``model_output`` is what ``model(rgb, depth)``
    returns in training mode (``((out, out8, out16, out32), loss_flop)``).  ``buckets`` is an optional
    :class:`dynmm_b200.dist.GradBuckets` whose all-reduce is captured between backward and the optimizer
    step.  Inputs are copied into static buffers, so shapes are fixed at construction."""

    def __init__(self, model, optimizer, loss_fn: Callable, rgb: Tensor, depth: Tensor, target: Tensor,
                 buckets=None, autocast: Optional[torch.dtype] = None, warmup: int = 3):
        if getattr(model, "ini_stage", False):
            raise RuntimeError("ini_stage draws branches on the host (model_skip_mod_globalgate.py:267-270): "
                               "not capturable")
        self.model, self.optimizer, self.loss_fn, self.buckets, self.autocast = model, optimizer, loss_fn, buckets, autocast
        self.rgb, self.depth, self.target = rgb.clone(), depth.clone(), target.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):           # allocates grads / momentum buffers / workspaces, binds contexts
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        # packed-weight caches created during capture live in the graph's pool and are refreshed by the
        # captured pack kernels only; drop the Python-side handles so eager calls re-pack
        for p in model.parameters():
            if hasattr(p, "_dynmm_pack"):
                del p._dynmm_pack

    def _step(self) -> Tensor:
        attached = self.buckets is not None and getattr(self.buckets, "_attached", False)
        if attached:
            self.buckets.zero()               # gradients are views into the flat buckets: one memset per bucket
        else:
            self.optimizer.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=self.autocast or torch.bfloat16, enabled=self.autocast is not None):
            out = self.model(self.rgb, self.depth)
        loss = self.loss_fn(out, self.target)
        loss.backward()
        if self.buckets is not None:
            # attached buckets: the hooks already launched every bucket's all-reduce during backward (overlapped);
            # this only makes the stream wait for them.  Otherwise: the simple post-backward exchange.
            self.buckets.allreduce(average=True)
        self.optimizer.step()
        return loss.detach()

    def __call__(self, rgb: Tensor, depth: Tensor, target: Tensor) -> Tensor:
        self.rgb.copy_(rgb, non_blocking=True)
        self.depth.copy_(depth, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss
