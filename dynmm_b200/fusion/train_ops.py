"""Differentiable convolution on the tcgen05 kernels.

The reference has no backward code of its own: ``loss.backward()`` (train.py:323) lets autograd
differentiate every ``F.conv2d`` of the encoder / decoder blocks (resnet.py:124-147, 66-84, 173-192;
model_utils.py:11-23).  Here the three convolution passes of a training step run on our kernels:

* forward        ``dynmm_conv_igemm_fwd``  (bias fused)
* data gradient  ``dynmm_conv_igemm_fwd`` on dy with mirrored taps / swapped channel roles; stride-2
  layers feed it the zero-interleaved dy
* weight gradient ``dynmm_conv_wgrad``     (pixel-K tcgen05 GEMM, deterministic split-K)

Tensors are bf16 ``channels_last`` (= the kernels' NHWC); master weights stay fp32 and are packed to
bf16 once per optimizer step (cached on the parameter, keyed by its version counter).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .. import ops

Tensor = torch.Tensor


def eligible(x: Tensor, weight: Tensor, stride, padding, dilation, groups) -> bool:
    """Can this convolution run on the tensor-core kernels?  (bf16 CUDA input, dense 'same'-padded
    1x1 / 1x3 / 3x1 / 3x3 taps, stride 1 or 2, channel counts in multiples of 8.)"""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    co, ci, kh, kw = weight.shape
    if groups != 1 or tuple(dilation) != (1, 1) or kh * kw > 9:
        return False
    if ci % 8 or co % 8 or ci < 16:
        return False
    if any(s not in (1, 2) for s in stride):
        return False
    if kh != 2 * padding[0] + 1 or kw != 2 * padding[1] + 1:
        return False
    return x.shape[2] >= 2 and x.shape[3] >= 2


def _packed(weight: Tensor) -> Tuple[Tensor, Tensor]:
    """(forward, data-gradient) packed bf16 copies of an fp32 master weight."""
    ver = weight._version
    cache = getattr(weight, "_dynmm_pack", None)
    if cache is None or cache[0] != ver or cache[1].device != weight.device:
        with torch.no_grad():
            cache = (ver,) + ops.pack_conv_weight_pair(weight)
        weight._dynmm_pack = cache
    return cache[1], cache[2]


def _nhwc(x: Tensor) -> Tensor:
    """[B,C,H,W] (any strides) -> contiguous NHWC view/copy."""
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], stride, padding):
        xn = _nhwc(x)
        co, ci, kh, kw = weight.shape
        wf, _ = _packed(weight)
        shift = bias.detach().float().contiguous() if bias is not None else None
        b, h_in, w_in, _ = xn.shape
        h_out = (h_in + 2 * padding[0] - kh) // stride[0] + 1
        w_out = (w_in + 2 * padding[1] - kw) // stride[1] + 1
        # the result is allocated in its final (channels_last) form and the kernel writes through an NHWC
        # view of it: returning a view from a custom Function would forbid the in-place ReLU that follows
        out = torch.empty((b, co, h_out, w_out), dtype=torch.bfloat16, device=x.device,
                          memory_format=torch.channels_last)
        ops.conv(xn, wf, c_out=co, kh=kh, kw=kw, stride=stride, pad=padding, shift=shift,
                 out=out.permute(0, 2, 3, 1), volatile_weights=True)
        ctx.save_for_backward(xn, weight)
        ctx.geom = (tuple(stride), tuple(padding), bias is not None)
        return out

    @staticmethod
    def backward(ctx, gy: Tensor):
        xn, weight = ctx.saved_tensors
        stride, padding, has_bias = ctx.geom
        co, ci, kh, kw = weight.shape
        gyn = _nhwc(gy.to(torch.bfloat16))
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            _, wd = _packed(weight)
            b, h_in, w_in, _ = xn.shape
            if stride != (1, 1):
                # zero-interleave dy back onto the input lattice; the unit-stride kernel does the rest
                up = torch.zeros(b, h_in, w_in, co, dtype=torch.bfloat16, device=gyn.device)
                up[:, 0:stride[0] * gyn.shape[1]:stride[0], 0:stride[1] * gyn.shape[2]:stride[1]] = gyn
            else:
                up = gyn
            gx = torch.empty((b, ci, h_in, w_in), dtype=torch.bfloat16, device=gyn.device,
                             memory_format=torch.channels_last)
            ops.conv(up, wd, c_out=ci, kh=kh, kw=kw, stride=(1, 1), pad=(kh - 1 - padding[0], kw - 1 - padding[1]),
                     out=gx.permute(0, 2, 3, 1), volatile_weights=True)
        if ctx.needs_input_grad[1]:
            gw = ops.conv_wgrad(xn, gyn, kh=kh, kw=kw, stride=stride, pad=padding, c_in=ci, c_out=co)
            gw = gw.to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            gb = ops.channel_sum(gyn, c=co)
        return gx, gw, gb, None, None


def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride, padding, dilation=(1, 1), groups=1) -> Tensor:
    """``F.conv2d`` with the tcgen05 forward/backward where :func:`eligible`, else the library conv
    (depthwise up-sampling stencils, the 1x1-on-1x1 SE layers) in the activation dtype."""
    if eligible(x, weight, stride, padding, dilation, groups):
        return _Conv2dFn.apply(x, weight, bias, tuple(stride), tuple(padding))
    if weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
        bias = bias.to(x.dtype) if bias is not None else None
    return F.conv2d(x, weight, bias, stride, padding, dilation, groups)
