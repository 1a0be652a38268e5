"""Robustness sweep of the gated RGB-D path (BASELINE.json configs[4]): Gaussian noise on RGB / depth, the gate's
branch distribution and images/s per noise level.

Host-side mirror of the evaluation loop of the reference (FusionDynMM/eval.py):

* ``set_seed(r)`` per run (eval.py:20-23, 78-79) -- python ``random``, numpy and torch generators;
* per BATCH one ``random.random()`` draw decides whether the batch is perturbed (eval.py:91-102):
  mode 0 -> RGB with probability 0.33, mode 1 -> depth with probability 0.33, mode 2 -> RGB for ``rand < 0.33``,
  depth for ``0.33 <= rand < 0.66``; ``x + noise * abs(x).mean() * randn_like(x)`` with the mean over the WHOLE batch
  tensor; mode -1 -> untouched;
* ``model(image, depth, True)`` (eval.py:109-115) under ``start_weight()`` / ``end_weight()`` so the per-sample branch
  choices are collected (model_skip_mod_globalgate.py:230-253, 273-274).

The noise is drawn with ``torch.randn_like`` on the tensor's own device, exactly like the reference, so on the same
device and seed both see the same perturbed inputs (the parity tests compare against the statement above on CPU).
The forward is the CUDA engine (``SkipGateESANet.forward`` / ``predict_labels``); nothing here falls back to a CPU
model.  Across ranks the sweep shards BATCHES round-robin (independent samples, no data-path collective); only the
5-bin histogram and the image / time totals are all-reduced (``dynmm_b200.dist.allreduce_histogram``).
"""
from __future__ import annotations

import random
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

MODE_NONE, MODE_RGB, MODE_DEPTH, MODE_BOTH = -1, 0, 1, 2


def set_seed(seed: int) -> None:
    """eval.py:20-23."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def add_noise(x: Tensor, noise: float) -> Tensor:
    """``x + noise * abs(x).mean() * randn_like(x)`` (eval.py:94, 97, 100, 102)."""
    return x + noise * abs(x).mean() * torch.randn_like(x)


def perturb(image: Tensor, depth: Tensor, mode: int, noise: float, rand_val: float) -> Tuple[Tensor, Tensor, int]:
    """One batch of eval.py:91-102.  Returns (image, depth, which) with which = -1 (clean), 0 (RGB noised) or
    1 (depth noised).  ``rand_val`` is the caller's ``random.random()`` draw for this batch (drawn for every batch,
    whatever the mode: eval.py:91)."""
    if mode not in (MODE_NONE, MODE_RGB, MODE_DEPTH, MODE_BOTH):
        raise ValueError(f"mode must be -1, 0, 1 or 2, got {mode}")
    which = -1
    if mode == MODE_RGB:
        if rand_val < 0.33:
            which = 0
    elif mode == MODE_DEPTH:
        if rand_val < 0.33:
            which = 1
    elif mode == MODE_BOTH:
        if rand_val < 0.33:
            which = 0
        elif rand_val < 0.66:
            which = 1
    if which == 0:
        image = add_noise(image, noise)
    elif which == 1:
        depth = add_noise(depth, noise)
    return image, depth, which


@dataclass
class SweepPoint:
    """Result of one (mode, noise) setting, summed over runs and ranks."""
    mode: int
    noise: float
    runs: int
    images: int = 0
    noised_batches: int = 0
    batches: int = 0
    histogram: List[int] = field(default_factory=lambda: [0] * 5)
    seconds: float = 0.0                       # max over ranks of the device time spent in forwards
    depth_flop_g: Optional[float] = None       # mean depth-encoder GFLOP per image (reference table :219)
    total_flop_g: Optional[float] = None       # mean total GFLOP per image (reference table :220)
    saved_pct: Optional[float] = None          # 1 - total / total(branch 4)

    @property
    def images_per_s(self) -> float:
        return self.images / self.seconds if self.seconds > 0 else 0.0

    def as_dict(self) -> Dict:
        tot = max(sum(self.histogram), 1)
        return {"mode": self.mode, "noise": self.noise, "runs": self.runs, "images": self.images,
                "batches": self.batches, "noised_batches": self.noised_batches,
                "gate_branch_histogram": list(self.histogram),
                "gate_branch_fraction": [h / tot for h in self.histogram],
                "images_per_s": self.images_per_s, "seconds": self.seconds,
                "depth_encoder_gflop_per_image": self.depth_flop_g, "total_gflop_per_image": self.total_flop_g,
                "flop_saved_pct": self.saved_pct}


def branch_histogram(weight: Tensor) -> Tensor:
    """int64[5] counts of the per-sample branch (arg-max of the gate weights; one-hot under hard gates)."""
    return torch.bincount(weight.argmax(dim=1), minlength=5)[:5].to(torch.int64)


def flop_summary(hist: Sequence[int], depth_enc_flop: Sequence[float], total_flop: Sequence[float]):
    """What ``end_weight(print_flop=True)`` prints (model_skip_mod_globalgate.py:240-250): expected depth-encoder and
    total GFLOP per image under the observed branch distribution, plus the saving against branch 4."""
    tot = float(sum(hist))
    if tot == 0:
        return None, None, None
    frac = [h / tot for h in hist]
    d = sum(f * x for f, x in zip(frac, depth_enc_flop))
    t = sum(f * x for f, x in zip(frac, total_flop))
    return d, t, 100.0 * (1.0 - t / float(total_flop[4]))


def run_point(model, batches: Callable[[int], Iterable[Tuple[Tensor, Tensor]]], mode: int, noise: float,
              num_runs: int = 1, rank: int = 0, world: int = 1, labels_only: bool = False,
              on_batch: Optional[Callable[[int, int, Tensor, Tensor], None]] = None) -> SweepPoint:
    """One (mode, noise) setting: ``num_runs`` seeded passes (eval.py:77-79) over ``batches(run)``, an iterable of
    device-resident ``(image, depth)`` pairs.  EVERY rank walks the whole batch list and draws ``random.random()`` for
    every batch (so the perturbation pattern does not depend on the number of ranks) but only forwards the batches
    ``i % world == rank``.  ``on_batch(run, index, prediction, weight)`` sees each forwarded batch's output (logits, or
    uint8 labels with ``labels_only``) and gate weights [B,5], e.g. to update a confusion matrix.  Returns this rank's partial SweepPoint (call
    ``reduce_point`` to sum over ranks)."""
    pt = SweepPoint(mode=mode, noise=float(noise), runs=num_runs)
    dev_hist = None
    cuda = None
    t_events = []
    for r in range(num_runs):
        set_seed(r)
        for i, (image, depth) in enumerate(batches(r)):
            rand_val = random.random()                       # eval.py:91: drawn for every batch
            mine = (i % world) == rank
            if not mine:
                continue
            image, depth, which = perturb(image, depth, mode, noise, rand_val)
            if cuda is None:
                cuda = image.is_cuda
            if cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            else:
                t0 = time.perf_counter()
            with torch.no_grad():
                if labels_only:
                    pred, weight = model.predict_labels(image, depth, return_weight=True)
                else:
                    pred, weight = model(image, depth, True, True)
            if cuda:
                e1.record()
                t_events.append((e0, e1))
            else:
                pt.seconds += time.perf_counter() - t0
            h = branch_histogram(weight)
            dev_hist = h if dev_hist is None else dev_hist + h
            pt.images += image.shape[0]
            pt.batches += 1
            pt.noised_batches += int(which >= 0)
            if on_batch is not None:
                on_batch(r, i, pred, weight)
    if cuda:
        torch.cuda.synchronize()
        pt.seconds = sum(a.elapsed_time(b) for a, b in t_events) * 1e-3
    if dev_hist is not None:
        pt.histogram = [int(v) for v in dev_hist.tolist()]
    return pt


def reduce_point(pt: SweepPoint, model=None, device=None) -> SweepPoint:
    """Sum a partial SweepPoint over ranks (histogram, images, batches: SUM; device seconds: MAX) and attach the
    reference's FLOP summary from the model's tables."""
    import torch.distributed as dist
    from .. import dist as ddp
    if dist.is_initialized() and dist.get_world_size() > 1:
        dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl"
                         else torch.device("cpu"))
        hist = torch.tensor(pt.histogram + [pt.images, pt.batches, pt.noised_batches], dtype=torch.int64, device=dev)
        ddp.allreduce_histogram(hist)
        secs = torch.tensor([pt.seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        vals = hist.tolist()
        pt.histogram, pt.images, pt.batches, pt.noised_batches = vals[:5], vals[5], vals[6], vals[7]
        pt.seconds = float(secs.item())
    if model is not None:
        pt.depth_flop_g, pt.total_flop_g, pt.saved_pct = flop_summary(
            pt.histogram, [float(v) for v in model.depth_enc_flop.tolist()], [float(v) for v in model.total_flop.tolist()])
    return pt


def sweep(model, batches: Callable[[int], Iterable[Tuple[Tensor, Tensor]]], noises: Sequence[float] = (0.0, 0.3, 0.6, 1.0),
          mode: int = MODE_DEPTH, num_runs: int = 1, rank: int = 0, world: int = 1,
          labels_only: bool = False) -> List[SweepPoint]:
    """configs[4]: one SweepPoint per noise level (depth noise by default), reduced over ranks."""
    return [reduce_point(run_point(model, batches, mode, s, num_runs, rank, world, labels_only), model) for s in noises]
