"""Data parallelism for the gated path: one process per GPU, ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in CPU tests).

The path shards by independent samples.  Inference needs no data-path
collective (replicas; only the 5-bin gate histogram is summed).  Training has
exactly one exchange step per optimizer step: a gradient all-reduce.  Gradients
are packed into flat, reverse-order buckets so each bucket's all-reduce can be
launched asynchronously while the rest of the backward is still running.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard(tensor: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """Contiguous batch shard of this rank (sizes differ by at most one)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b = tensor.shape[0]
    lo = (b * rank) // world
    hi = (b * (rank + 1)) // world
    return tensor[lo:hi]


class GradBuckets:
    """Flat gradient buckets over the trainable parameters, built in REVERSE
    registration order (the order backward produces gradients)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)

    def allreduce(self, average: bool = True) -> None:
        """All-reduce every bucket (async launches, one wait at the end) and scatter the
        result back into ``p.grad``.  Parameters without a gradient contribute zeros."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        works = []
        for i, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            ref = bucket[0]
            flat = self._flat[i]
            if flat is None or flat.numel() != n or flat.device != ref.device:
                flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=ref.device)
            off = 0
            for p in bucket:
                g = p.grad
                view = flat[off:off + p.numel()]
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g.reshape(-1))
                off += p.numel()
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
        for i, (bucket, work) in enumerate(zip(self.buckets, works)):
            work.wait()
            flat = self._flat[i]
            if average:
                flat.div_(world)
            off = 0
            for p in bucket:
                view = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = view.clone()
                else:
                    p.grad.copy_(view)
                off += p.numel()


def allreduce_histogram(hist: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank gate-branch histogram (int64[5]) -- the only collective of the eval path."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Make replicas identical (parameters and buffers, e.g. BN running statistics)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)       # in-place on the tensor itself: bumps its version counter (engine cache key)
    if hasattr(module, "invalidate_engine"):
        module.invalidate_engine()
