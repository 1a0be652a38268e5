"""Inference engine of the gated RGB-D encoder/decoder on the CUDA kernels.

Weight preparation (BN folding, bf16 K-major repacking) happens once per
``state_dict``; the forward is a fixed sequence of C-ABI launches whose
data-dependent parts (which samples run which depth stage) are resolved on the
device, so the whole forward can be captured in a CUDA graph.
"""
from __future__ import annotations

from typing import Dict

import torch

from .. import ops

Tensor = torch.Tensor


def pack_gate(sd: Dict[str, Tensor], prefix: str = "gate_layer.") -> Dict[str, Tensor]:
    """GlobalGate parameters (model_skip_mod_globalgate.py:379-386) in the layout
    dynmm_global_gate_logits expects; conv bias + eval BN folded to scale/shift."""
    g = lambda k: sd[prefix + k].detach().float()
    s1, b1 = ops.fold_bn(g("conv.1.weight"), g("conv.1.bias"), g("conv.1.running_mean"), g("conv.1.running_var"),
                         1e-5, g("conv.0.bias"))
    s2, b2 = ops.fold_bn(g("conv.4.weight"), g("conv.4.bias"), g("conv.4.running_mean"), g("conv.4.running_var"),
                         1e-5, g("conv.3.bias"))
    return {
        "w1": g("conv.0.weight").permute(0, 2, 3, 1).contiguous(),   # [8][5][5][128]
        "s1": s1, "b1": b1,
        "w2": g("conv.3.weight").permute(0, 2, 3, 1).contiguous(),   # [8][5][5][8]
        "s2": s2, "b2": b2,
        "wfc": g("fc.weight").reshape(g("fc.weight").shape[0], -1).contiguous(),
    }
