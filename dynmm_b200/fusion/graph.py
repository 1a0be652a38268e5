"""CUDA-graph replay of the engine's launch sequence.

The forward is ~170 dependent launches of a few microseconds each on two
streams; replaying one captured graph removes the per-launch host cost.  Gate
decisions stay on the device (kernels read counts / slot maps from memory), so
ONE graph serves every outcome of the gate.
"""
from __future__ import annotations

import torch


class GraphedForward:
    def __init__(self, engine, rgb, depth, modes, labels_only: bool = False):
        self.rgb = torch.empty_like(rgb, memory_format=torch.contiguous_format)
        self.depth = torch.empty_like(depth, memory_format=torch.contiguous_format)
        self.rgb.copy_(rgb)
        self.depth.copy_(depth)
        extra = {}
        self.labels = None
        if labels_only:
            b, _, h, w = rgb.shape
            self.labels = torch.empty(b, h, w, dtype=torch.uint8, device=rgb.device)
            extra = dict(labels=self.labels, want_logits=False)
        cur = torch.cuda.current_stream()
        warm = torch.cuda.Stream(device=rgb.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(2):
                engine.forward(self.rgb, self.depth, **modes, **extra)
        cur.wait_stream(warm)
        torch.cuda.synchronize(rgb.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, self.weight = engine.forward(self.rgb, self.depth, **modes, **extra)
        self.launches = engine.launches
        # program images are re-uploaded from pinned host memory on every replay: keep them (and their tensors) alive
        self._programs = list(engine.programs)

    def __call__(self, rgb, depth):
        """Outputs are static buffers, valid until the next call."""
        self.rgb.copy_(rgb, non_blocking=True)
        self.depth.copy_(depth, non_blocking=True)
        self.graph.replay()
        return (self.labels if self.labels is not None else self.out), self.weight


class GraphedLocalForward:
    """The same for ``FusionEngine.forward_local`` (local-gate SkipESANet): every stage's slot order is recomputed on
    the device from the gate weights and the Gumbel noise comes from the device generator (graph-safe Philox offsets),
    so one captured graph serves every decision.  Not for ``random_policy`` (CPU ``torch.randint``)."""

    def __init__(self, engine, rgb, depth, modes):
        self.rgb = torch.empty_like(rgb, memory_format=torch.contiguous_format)
        self.depth = torch.empty_like(depth, memory_format=torch.contiguous_format)
        self.rgb.copy_(rgb)
        self.depth.copy_(depth)
        cur = torch.cuda.current_stream()
        warm = torch.cuda.Stream(device=rgb.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(2):
                engine.forward_local(self.rgb, self.depth, **modes)
        cur.wait_stream(warm)
        torch.cuda.synchronize(rgb.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, self.weights, self.counts = engine.forward_local(self.rgb, self.depth, **modes)
        self.launches = engine.launches

    def __call__(self, rgb, depth):
        """Outputs are static buffers, valid until the next call."""
        self.rgb.copy_(rgb, non_blocking=True)
        self.depth.copy_(depth, non_blocking=True)
        self.graph.replay()
        return self.out, self.weights, self.counts
