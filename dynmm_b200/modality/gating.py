"""Gate + mixture arithmetic shared by the modality-level DynMM nets.

* ``DiffSoftmax``: same function as in fusion-level DynMM (custom CUDA forward
  and backward on the GPU).
* ``mix``: ``sum_e w[:,e] * pred_e`` through ``dynmm_softgate_mix_fwd/bwd``.
* ``routed_mix``: hard gates in inference -- each expert is evaluated ONLY on the
  rows routed to it (stable compaction on the device), which is the computation
  saving the reference only accounts for on paper (imdb_dyn.py:66,96).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from .. import ops
from ..fusion.autograd_ops import diff_softmax


def DiffSoftmax(logits, tau=1.0, hard=False, dim=-1):
    """imdb_dyn.py:16-26 / affect_dyn.py:18-28."""
    return diff_softmax(logits, tau, hard, dim)


class _MixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, *preds):
        w = w.contiguous()
        preds = [p.contiguous() for p in preds]
        ctx.save_for_backward(w, *preds)
        return ops.softgate_mix_fwd(preds, w)

    @staticmethod
    def backward(ctx, grad):
        w, *preds = ctx.saved_tensors
        grads, gw = ops.softgate_mix_bwd(grad, preds, w, list(ctx.needs_input_grad[1:]))
        return (gw if ctx.needs_input_grad[0] else None, *grads)


def mix(weight: torch.Tensor, preds: Sequence[torch.Tensor]) -> torch.Tensor:
    """output = sum_e weight[:, e:e+1] * preds[e]  (imdb_dyn.py:100, affect_dyn.py:95,164)."""
    if weight.is_cuda and weight.dtype == torch.float32 and all(p.dtype == torch.float32 and p.dim() == 2 for p in preds) \
            and len(preds) <= 4:
        return _MixFn.apply(weight, *preds)
    out = weight[:, 0:1] * preds[0]
    for e in range(1, len(preds)):
        out = out + weight[:, e:e + 1] * preds[e]
    return out


def can_route(weight: torch.Tensor, hard: bool, training: bool) -> bool:
    return hard and not training and weight.is_cuda and not torch.is_grad_enabled()


def routed_mix(weight: torch.Tensor, experts: Sequence[Callable[[torch.Tensor], torch.Tensor]], out_dim: int):
    """weight [B,E] one-hot.  ``experts[e](rows)`` evaluates expert e on the int64 row ids
    ``rows`` and returns [len(rows), out_dim].  One host sync (the row counts size the experts'
    GEMMs); experts with no rows are not launched at all."""
    b, ne = weight.shape
    plans = [ops.compact_rows(weight, e) for e in range(ne)]
    counts = torch.cat([p[2] for p in plans]).tolist()           # the single device->host sync
    preds, rows = [], []
    for e, (idx, inv, _) in enumerate(plans):
        k = counts[e]
        if k:
            preds.append(experts[e](idx[:k].long()).float().contiguous())
        else:
            preds.append(torch.zeros(1, out_dim, device=weight.device))
        rows.append(inv)
    return ops.softgate_mix_fwd(preds, weight.contiguous(), rows=rows), counts
