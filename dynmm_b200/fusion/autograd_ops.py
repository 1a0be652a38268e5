"""Differentiable gate ops: custom CUDA forward/backward where the tensors live
on the GPU, the reference's PyTorch formulation otherwise (CPU tensors)."""
from __future__ import annotations

import torch

from .. import ops


class _DiffSoftmaxFn(torch.autograd.Function):
    """dynmm_diffsoftmax_fwd / _bwd.  Hard or soft, the gradient is the tempered
    softmax Jacobian (straight-through; model_skip_mod_globalgate.py:22-26)."""

    @staticmethod
    def forward(ctx, logits, tau, hard):
        y, y_soft, _ = ops.diffsoftmax_fwd(logits.contiguous(), tau, hard)
        ctx.save_for_backward(y_soft)
        ctx.tau = tau
        return y

    @staticmethod
    def backward(ctx, grad):
        (y_soft,) = ctx.saved_tensors
        return ops.diffsoftmax_bwd(grad, y_soft, ctx.tau), None, None


def diff_softmax(logits, tau=1.0, hard=False, dim=-1):
    if logits.dim() > 2 and dim in (1, -3) and logits.shape[2:].numel() == 1:
        # [B,n,1,1] as produced by the gate's 1x1 conv
        return diff_softmax(logits.flatten(1), tau, hard, 1).view_as(logits)
    if logits.is_cuda and logits.dim() == 2 and dim in (-1, 1) and logits.dtype == torch.float32 \
            and logits.shape[1] <= 32:
        return _DiffSoftmaxFn.apply(logits, float(tau), bool(hard))
    y_soft = (logits / tau).softmax(dim)
    if not hard:
        return y_soft
    index = y_soft.max(dim, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(dim, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


class _GatedBlendFn(torch.autograd.Function):
    """fuse = rgb + g[n] * depth  ==  w*rgb + (1-w)*(rgb+depth) with g = 1-w
    (model_skip_mod_globalgate.py:279-283).  grad_g[n] = <grad, depth_n>."""

    @staticmethod
    def forward(ctx, rgb, depth, g):
        rgb, depth, g = rgb.contiguous(), depth.contiguous(), g.contiguous()
        ctx.save_for_backward(depth, g)
        return ops.gated_add_f32_fwd(rgb, depth, g)

    @staticmethod
    def backward(ctx, grad):
        depth, g = ctx.saved_tensors
        need_d = ctx.needs_input_grad[1]
        grad_d, grad_g = ops.gated_add_f32_bwd(grad, depth, g, need_grad_b=need_d)
        return (grad if ctx.needs_input_grad[0] else None), grad_d, (grad_g if ctx.needs_input_grad[2] else None)


def gated_blend(rgb, depth, g):
    if rgb.is_cuda and rgb.dtype == torch.float32 and (rgb.numel() // rgb.shape[0]) % 4 == 0:
        return _GatedBlendFn.apply(rgb, depth, g.float())
    return rgb + g.view(-1, 1, 1, 1) * depth
