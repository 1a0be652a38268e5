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
    """Gradient exchange of the data-parallel training step (the reference is single-GPU: train.py:299-324 has no
    exchange; this is the ONE collective north_star names).

    Parameters are grouped into flat fp32 buckets in REVERSE registration order -- the order backward produces
    gradients.  :meth:`attach` makes every ``p.grad`` a VIEW into its bucket (autograd then accumulates straight into
    the bucket: no gather / scatter copies) and registers post-accumulate hooks: the moment the last gradient of a
    bucket has been written, that bucket's all-reduce is launched asynchronously, so the exchange of the late layers
    overlaps the backward of the early ones.  :meth:`finish` (before ``optimizer.step()``) waits for the outstanding
    reductions; the average is taken by NCCL itself (``ReduceOp.AVG``) or by one in-place divide on gloo.

    ``allreduce()`` without ``attach`` keeps the simple post-backward form (used by the gloo test and eager tools):
    gradients are copied into the buckets, reduced, and copied back."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._attached = False
        self._hooks: list = []
        self._pending: List[int] = []
        self._works: list = []
        self._bucket_of = {}
        self.launched_in_backward = 0          # buckets whose all-reduce was started from a hook in the last step
        self.enabled = True                    # False: no collective at all (measuring the step without its exchange)

    # ------------------------------------------------------------------ helpers
    def _flat_for(self, i: int) -> torch.Tensor:
        bucket = self.buckets[i]
        n = sum(p.numel() for p in bucket)
        ref = bucket[0]
        flat = self._flat[i]
        if flat is None or flat.numel() != n or flat.device != ref.device:
            flat = self._flat[i] = torch.zeros(n, dtype=torch.float32, device=ref.device)
        return flat

    def _active(self) -> bool:
        return self.enabled and dist.is_initialized() and dist.get_world_size() > 1

    def _launch(self, i: int):
        flat = self._flat[i]
        if dist.get_backend() == "nccl":
            self._works.append((i, dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True), False))
        else:
            self._works.append((i, dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), True))

    # ------------------------------------------------------------------ overlapped form
    def attach(self) -> "GradBuckets":
        """Gradients become views into the buckets; hooks launch each bucket's all-reduce during backward."""
        if self._attached:
            return self
        for i, bucket in enumerate(self.buckets):
            flat = self._flat_for(i)
            off = 0
            for p in bucket:
                if p.dtype != torch.float32:
                    raise TypeError("GradBuckets.attach needs fp32 master parameters")
                p.grad = flat[off:off + p.numel()].view_as(p)
                self._bucket_of[p] = i
                off += p.numel()
        self._pending = [len(b) for b in self.buckets]
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._attached = True
        return self

    def detach(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self._attached = False

    def _on_grad(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self._active():
            self._launch(i)
            self.launched_in_backward += 1

    def zero(self):
        """Zero the gradient buckets (one memset per bucket) and re-arm the hooks; call instead of
        ``optimizer.zero_grad()`` at the top of a step (``zero_grad(set_to_none=False)`` works too, call this after)."""
        for i in range(len(self.buckets)):
            self._flat_for(i).zero_()
        self._pending = [len(b) for b in self.buckets]
        self._works = []
        self.launched_in_backward = 0

    def finish(self):
        """Wait (stream-wise) for the reductions started during backward; buckets whose hooks did not all fire
        (parameters that received no gradient this step) are reduced now."""
        if not self._active():
            return
        started = {i for i, _, _ in self._works}
        for i in range(len(self.buckets)):
            if i not in started:
                self._launch(i)
        world = dist.get_world_size()
        for i, work, need_div in self._works:
            work.wait()
            if need_div:
                self._flat[i].div_(world)
        self._works = []

    # ------------------------------------------------------------------ simple form
    def allreduce(self, average: bool = True) -> None:
        """Post-backward exchange.  Attached: equivalent to :meth:`finish`.  Otherwise: copy ``p.grad`` into the
        buckets, all-reduce every bucket (async launches, one wait at the end), copy back.  Parameters without a
        gradient contribute zeros."""
        if not self._active():
            return
        if self._attached:
            self.finish()
            return
        world = dist.get_world_size()
        works = []
        for i, bucket in enumerate(self.buckets):
            flat = self._flat_for(i)
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
