"""Drop-in ``nn.Module`` surface of fusion-level DynMM.

Same constructor signatures, attribute names, ``state_dict`` keys and return
conventions as the reference (``FusionDynMM/src/models`` of zihuixue/DynMM;
file:line citations below are relative to that directory), so the reference's
``train.py`` / ``eval.py`` / ``build_model`` work unchanged and checkpoints load
with ``strict=True``.

Two bodies behind that surface:

* eval mode on a CUDA device -> :class:`~dynmm_b200.fusion.engine.FusionEngine`
  (hand-written sm_100a kernels through the C ABI; gated-off depth stages are
  really skipped).  No fallback: a missing library raises.
* training mode (and CPU tensors) -> the differentiable PyTorch graph below,
  which needs every sample's depth features for the gate gradient and batch
  statistics for BatchNorm (hard part 1).  On CUDA its gate
  ops (DiffSoftmax, gated blend) run on the custom kernels with custom backward.
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence

import numpy as np
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .autograd_ops import diff_softmax, gated_blend

Tensor = torch.Tensor

DEPTH_ENC_FLOP_R34 = [0.2506752, 3.1113216, 6.9470208, 12.66432, 15.538944]
TOTAL_FLOP_R34 = [22.37101509, 25.23166149, 29.06736069, 34.78465989, 37.65928389]
DEPTH_ENC_FLOP_OTHER = [0.2506752, 4.39420573, 10.72382115, 19.71582947, 24.679084]
TOTAL_FLOP_OTHER = [32.5854654, 36.728995928, 43.058611352, 52.050619672, 57.0138742]


def DiffSoftmax(logits, tau=1.0, hard=False, dim=-1):
    """model_skip_mod_globalgate.py:20-30."""
    return diff_softmax(logits, tau, hard, dim)


class Swish(nn.Module):          # model_utils.py:100-106
    def forward(self, x):
        return x * torch.sigmoid(x)


class Hswish(nn.Module):         # model_utils.py:109-115
    def __init__(self, inplace=True):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return x * F.relu6(x + 3.0, inplace=self.inplace) / 6.0


def _make_activation(name: str) -> nn.Module:
    name = name.lower()
    if name == "relu":
        return nn.ReLU(inplace=True)
    if name in ("swish", "silu"):
        return Swish()
    if name == "hswish":
        return Hswish()
    raise NotImplementedError(
        'Only relu, swish and hswish as activation function are supported so far. Got {}'.format(name))


class Conv2d(nn.Conv2d):
    """``nn.Conv2d`` (same parameters / state_dict keys) whose bf16 CUDA inputs run on the tcgen05
    forward, data-gradient and weight-gradient kernels (train_ops.py); fp32 inputs take the
    stock library path, i.e. the reference's arithmetic."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.bfloat16 and not torch.is_autocast_enabled():
            from .train_ops import conv2d
            return conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return super().forward(x)


class ConvBNAct(nn.Sequential):
    """model_utils.py:11-23 (keys: conv, bn, act)."""

    def __init__(self, channels_in, channels_out, kernel_size, activation=nn.ReLU(inplace=True), dilation=1, stride=1):
        super().__init__()
        self.add_module("conv", Conv2d(channels_in, channels_out, kernel_size, stride=stride,
                                          padding=kernel_size // 2 + dilation - 1, dilation=dilation, bias=False))
        self.add_module("bn", nn.BatchNorm2d(channels_out))
        self.add_module("act", activation)


class SqueezeAndExcitation(nn.Module):
    """model_utils.py:36-51 (keys: fc.0, fc.2)."""

    def __init__(self, channel, reduction=16, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.fc = nn.Sequential(Conv2d(channel, channel // reduction, 1), activation,
                                Conv2d(channel // reduction, channel, 1), nn.Sigmoid())

    def forward(self, x):
        return x * self.fc(F.adaptive_avg_pool2d(x, 1))


class SqueezeAndExciteFusionAdd(nn.Module):
    """rgb_depth_fusion.py:13-26."""

    def __init__(self, channels_in, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.se_rgb = SqueezeAndExcitation(channels_in, activation=activation)
        self.se_depth = SqueezeAndExcitation(channels_in, activation=activation)

    def forward(self, rgb, depth):
        return self.se_rgb(rgb) + self.se_depth(depth)


# ------------------------------------------------------------------ encoder blocks

class NonBottleneck1D(nn.Module):
    """resnet.py:87-147 -- factorised residual block; BN eps 1e-3."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.conv3x1_1 = Conv2d(inplanes, planes, (3, 1), stride=(stride, 1), padding=(1, 0), bias=True)
        self.conv1x3_1 = Conv2d(planes, planes, (1, 3), stride=(1, stride), padding=(0, 1), bias=True)
        self.bn1 = nn.BatchNorm2d(planes, eps=1e-3)
        self.act = activation
        self.conv3x1_2 = Conv2d(planes, planes, (3, 1), padding=(dilation, 0), dilation=(dilation, 1), bias=True)
        self.conv1x3_2 = Conv2d(planes, planes, (1, 3), padding=(0, dilation), dilation=(1, dilation), bias=True)
        self.bn2 = nn.BatchNorm2d(planes, eps=1e-3)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.act(self.conv3x1_1(x))
        y = self.act(self.bn1(self.conv1x3_1(y)))
        y = self.act(self.conv3x1_2(y))
        y = self.bn2(self.conv1x3_2(y))
        idn = x if self.downsample is None else self.downsample(x)
        return self.act(y + idn)


class BasicBlock(nn.Module):
    """resnet.py:42-84."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride, dilation, dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.act = activation
        self.conv2 = Conv2d(planes, planes, 3, 1, dilation, dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.act(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        idn = x if self.downsample is None else self.downsample(x)
        return self.act(y + idn)


class Bottleneck(nn.Module):
    """resnet.py:150-192."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, dilation, dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.act = activation
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.act(self.bn1(self.conv1(x)))
        y = self.act(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        idn = x if self.downsample is None else self.downsample(x)
        return self.act(y + idn)


_BLOCKS = {"NonBottleneck1D": NonBottleneck1D, "BasicBlock": BasicBlock, "Bottleneck": Bottleneck}
_DEPTHS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}


class ResNet(nn.Module):
    """Staged encoder (resnet.py:195-379): conv1/bn1 stem and layer1..4; the caller
    applies the max-pool between ``forward_first_conv`` and ``forward_layer1``."""

    def __init__(self, layers: Sequence[int], block, input_channels=3, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.conv1 = Conv2d(input_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.act = activation
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        e = block.expansion
        self.down_2_channels_out = 64
        self.down_4_channels_out, self.down_8_channels_out = 64 * e, 128 * e
        self.down_16_channels_out, self.down_32_channels_out = 256 * e, 512 * e
        inplanes = 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            blocks = []
            for b in range(n):
                ds = None
                if b == 0 and (stride != 1 or inplanes != planes * e):
                    ds = nn.Sequential(Conv2d(inplanes, planes * e, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * e))
                blocks.append(block(inplanes, planes, stride if b == 0 else 1, ds, activation=activation))
                inplanes = planes * e
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        for m in self.modules():          # resnet.py:264-270
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward_first_conv(self, x):
        return self.act(self.bn1(self.conv1(x)))

    def forward_layer1(self, x):
        return self.layer1(x)

    def forward_layer2(self, x):
        return self.layer2(x)

    def forward_layer3(self, x):
        return self.layer3(x)

    def forward_layer4(self, x):
        return self.layer4(x)

    def forward(self, x):
        x = self.maxpool(self.forward_first_conv(x))
        f1 = self.layer1(x)
        f2 = self.layer2(f1)
        f3 = self.layer3(f2)
        return [self.layer4(f3), f3, f2, f1]


def _build_encoder(name: str, block: str, input_channels: int, activation, pretrained_on_imagenet: bool,
                   which: str) -> ResNet:
    if name not in _DEPTHS:
        raise NotImplementedError('Only ResNets are supported for {}. Got {}'.format(which, name))
    if pretrained_on_imagenet:
        raise NotImplementedError(
            "pretrained_on_imagenet needs network access / external checkpoints; load weights with "
            "load_state_dict (keys are identical to the reference's) instead")
    if name == "resnet50":
        blk = Bottleneck
    elif block in _BLOCKS:
        blk = _BLOCKS[block]
    else:
        raise NotImplementedError('Block {} is not implemented'.format(block))
    return ResNet(_DEPTHS[name], blk, input_channels, activation)


def ResNet18(block="BasicBlock", input_channels=3, activation=nn.ReLU(inplace=True), pretrained_on_imagenet=False,
             pretrained_dir=None):
    return _build_encoder("resnet18", block, input_channels, activation, pretrained_on_imagenet, "encoder")


def ResNet34(block="BasicBlock", input_channels=3, activation=nn.ReLU(inplace=True), pretrained_on_imagenet=False,
             pretrained_dir=None):
    return _build_encoder("resnet34", block, input_channels, activation, pretrained_on_imagenet, "encoder")


def ResNet50(input_channels=3, activation=nn.ReLU(inplace=True), pretrained_on_imagenet=False, **_):
    return _build_encoder("resnet50", "Bottleneck", input_channels, activation, pretrained_on_imagenet, "encoder")


# ------------------------------------------------------------------ context module / decoder

class PyramidPoolingModule(nn.Module):
    """context_modules.py:47-87 (keys: features.{i}.1.{conv,bn}, final_conv.{conv,bn})."""

    def __init__(self, in_dim, out_dim, bins=(1, 5), activation=nn.ReLU(inplace=True), upsampling_mode="nearest"):
        super().__init__()
        red = in_dim // len(bins)
        self.upsampling_mode = upsampling_mode
        self.features = nn.ModuleList(
            [nn.Sequential(nn.AdaptiveAvgPool2d(b), ConvBNAct(in_dim, red, 1, activation=activation)) for b in bins])
        self.final_conv = ConvBNAct(in_dim + red * len(bins), out_dim, 1, activation=activation)

    def forward(self, x):
        size = (int(x.shape[2]), int(x.shape[3]))
        outs = [x]
        for f in self.features:
            y = f(x)
            if self.upsampling_mode == "nearest":
                outs.append(F.interpolate(y, size, mode="nearest"))
            elif self.upsampling_mode == "bilinear":
                outs.append(F.interpolate(y, size, mode="bilinear", align_corners=False))
            else:
                raise NotImplementedError('For the PyramidPoolingModule only nearest and bilinear interpolation '
                                          f'are supported. Got: {self.upsampling_mode}')
        return self.final_conv(torch.cat(outs, 1))


def get_context_module(name, channels_in, channels_out, input_size, activation, upsampling_mode="bilinear"):
    """context_modules.py:16-44.  'appm*' (fixed-size adaptive variant) is not provided."""
    if "appm" in name:
        raise NotImplementedError("context module 'appm' is outside the accelerated path; use 'ppm'")
    if "ppm" in name:
        bins = (1, 2, 4, 8) if name == "ppm-1-2-4-8" else (1, 5)
        return PyramidPoolingModule(channels_in, channels_out, bins, activation, upsampling_mode), channels_out
    return nn.Identity(), channels_in


class Upsample(nn.Module):
    """model.py:360-410 (keys: conv.weight [C,1,3,3], conv.bias for the learned modes)."""

    def __init__(self, mode, channels=None):
        super().__init__()
        self.align_corners = False if mode == "bilinear" else None
        if "learned-3x3" in mode:
            if mode == "learned-3x3":
                self.pad = nn.ReplicationPad2d((1, 1, 1, 1))
                self.conv = Conv2d(channels, channels, 3, groups=channels, padding=0)
            else:
                self.pad = nn.Identity()
                self.conv = Conv2d(channels, channels, 3, groups=channels, padding=1)
            stencil = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]])
            with torch.no_grad():
                self.conv.weight.copy_(stencil.expand(channels, 1, 3, 3))
                self.conv.bias.zero_()
            self.mode = "nearest"
        else:
            self.pad, self.conv, self.mode = nn.Identity(), nn.Identity(), mode

    def forward(self, x):
        size = (int(x.shape[2] * 2), int(x.shape[3] * 2))
        x = F.interpolate(x, size, mode=self.mode, align_corners=self.align_corners)
        return self.conv(self.pad(x))


class DecoderModule(nn.Module):
    """model.py:311-357."""

    def __init__(self, channels_in, channels_dec, activation=nn.ReLU(inplace=True), nr_decoder_blocks=1,
                 encoder_decoder_fusion="add", upsampling_mode="bilinear", num_classes=37):
        super().__init__()
        self.upsampling_mode = upsampling_mode
        self.encoder_decoder_fusion = encoder_decoder_fusion
        self.conv3x3 = ConvBNAct(channels_in, channels_dec, 3, activation=activation)
        self.decoder_blocks = nn.Sequential(
            *[NonBottleneck1D(channels_dec, channels_dec, activation=activation) for _ in range(nr_decoder_blocks)])
        self.upsample = Upsample(upsampling_mode, channels_dec)
        self.side_output = Conv2d(channels_dec, num_classes, 1)

    def forward(self, decoder_features, encoder_features):
        out = self.decoder_blocks(self.conv3x3(decoder_features))
        side = self.side_output(out) if self.training else None
        out = self.upsample(out)
        if self.encoder_decoder_fusion == "add":
            out = out + encoder_features
        return out, side


class Decoder(nn.Module):
    """model.py:244-308: eval -> logits; train -> (logits, 1/8, 1/16, 1/32 side outputs)."""

    def __init__(self, channels_in, channels_decoder, activation=nn.ReLU(inplace=True), nr_decoder_blocks=(1, 1, 1),
                 encoder_decoder_fusion="add", upsampling_mode="bilinear", num_classes=37):
        super().__init__()
        cin = channels_in
        for i in range(3):
            setattr(self, f"decoder_module_{i + 1}",
                    DecoderModule(cin, channels_decoder[i], activation, nr_decoder_blocks[i], encoder_decoder_fusion,
                                  upsampling_mode, num_classes))
            cin = channels_decoder[i]
        self.conv_out = Conv2d(cin, num_classes, 3, padding=1)
        self.upsample1 = Upsample(upsampling_mode, num_classes)
        self.upsample2 = Upsample(upsampling_mode, num_classes)

    def forward(self, enc_outs):
        x, s16, s8, s4 = enc_outs
        x, side32 = self.decoder_module_1(x, s16)
        x, side16 = self.decoder_module_2(x, s8)
        x, side8 = self.decoder_module_3(x, s4)
        x = self.upsample2(self.upsample1(self.conv_out(x)))
        if self.training:
            return x, side8, side16, side32
        return x


# ------------------------------------------------------------------ gate

class GlobalGate(nn.Module):
    """model_skip_mod_globalgate.py:375-394 (keys: conv.{0,1,3,4}, fc)."""

    def __init__(self, branch_num, hidden_dim=8):
        super().__init__()
        self.bnum = branch_num
        self.conv = nn.Sequential(
            Conv2d(128, hidden_dim, 5, 2), nn.BatchNorm2d(hidden_dim), nn.Tanh(),
            Conv2d(hidden_dim, hidden_dim, 5, 2), nn.BatchNorm2d(hidden_dim), nn.Tanh())
        self.fc = Conv2d(hidden_dim, branch_num, 1, bias=False)

    def logits(self, rgb, depth):
        y = self.conv(torch.cat([rgb, depth], 1))
        return self.fc(F.adaptive_avg_pool2d(y, 1)).flatten(1)

    def forward(self, rgb, depth, temp=1.0, hard=False):
        return diff_softmax(self.logits(rgb, depth), temp, hard, 1)


# ------------------------------------------------------------------ the model

def stem_channels_last(encoder, x):
    """``forward_first_conv`` (resnet.py:352-358) for the bf16 training graph: the fp32 stem convolution as written, then
    ONE layout pass to channels_last so that BatchNorm, ReLU, the add and the max-pools (and their backward passes) run
    ATen's NHWC kernels instead of cuDNN's NCHW ones (measured at batch 8: BN backward 0.73 ms per stem map in NCHW).
    Values stay fp32 -- the gate's decisions must match the reference."""
    if os.environ.get("DYNMM_TRAIN_STEM") == "nchw":          # comparison switch
        return encoder.forward_first_conv(x)
    y = encoder.conv1(x).contiguous(memory_format=torch.channels_last)
    return encoder.act(encoder.bn1(y))


class SkipGateESANet(nn.Module):
    """Global-gate dynamic ESANet (model_skip_mod_globalgate.py:33-322)."""

    def __init__(self, height=480, width=640, num_classes=40, encoder_rgb="resnet34", encoder_depth="resnet34",
                 encoder_block="NonBottleneck1D", channels_decoder=None, pretrained_on_imagenet=False,
                 pretrained_dir="./trained_models/imagenet", activation="relu", encoder_decoder_fusion="add",
                 context_module="ppm", nr_decoder_blocks=None, fuse_depth_in_rgb_encoder="add",
                 upsampling="learned-3x3-zeropad", temp=1, block_rule=None):
        super().__init__()
        channels_decoder = [128, 128, 128] if channels_decoder is None else list(channels_decoder)
        nr_decoder_blocks = [3, 3, 3] if nr_decoder_blocks is None else list(nr_decoder_blocks)
        self.fuse_depth_in_rgb_encoder = fuse_depth_in_rgb_encoder
        self.block_rule = block_rule if block_rule else [1, 1, 1, 1]     # stored, unused (:62)
        self.activation = _make_activation(activation)
        if encoder_rgb == "resnet50" or encoder_depth == "resnet50":
            warnings.warn("Parameter encoder_block is ignored for ResNet50. ResNet50 always uses Bottleneck")
        self.encoder_rgb = _build_encoder(encoder_rgb, encoder_block, 3, self.activation, pretrained_on_imagenet,
                                          "encoder_rgb")
        self.encoder_depth = _build_encoder(encoder_depth, encoder_block, 1, self.activation, pretrained_on_imagenet,
                                            "encoder_depth")
        enc = self.encoder_rgb
        self.channels_decoder_in = enc.down_32_channels_out
        stage_ch = (enc.down_4_channels_out, enc.down_8_channels_out, enc.down_16_channels_out,
                    enc.down_32_channels_out)
        if fuse_depth_in_rgb_encoder == "SE-add":
            for i, c in enumerate((64,) + stage_ch):
                setattr(self, f"se_layer{i}", SqueezeAndExciteFusionAdd(c, activation=self.activation))
        if encoder_decoder_fusion == "add":
            for i, (c_enc, c_dec) in enumerate(zip(stage_ch[:3], channels_decoder[::-1])):
                layers = [ConvBNAct(c_enc, c_dec, 1, activation=self.activation)] if c_enc != c_dec else []
                setattr(self, f"skip_layer{i + 1}", nn.Sequential(*layers))
        elif encoder_decoder_fusion == "None":
            for i in range(4):
                setattr(self, f"skip_layer{i}", nn.Identity())
        if "learned-3x3" in upsampling:
            warnings.warn("for the context module the learned upsampling is not possible as the feature maps are "
                          "not upscaled by the factor 2. We will use nearest neighbor instead.")
            ctx_up = "nearest"
        else:
            ctx_up = upsampling
        self.context_module, ch_ctx = get_context_module(context_module, self.channels_decoder_in,
                                                         channels_decoder[0], (height // 32, width // 32),
                                                         self.activation, ctx_up)
        self.decoder = Decoder(ch_ctx, channels_decoder, self.activation, nr_decoder_blocks, encoder_decoder_fusion,
                               upsampling, num_classes)
        # gating network and the mutable mode attributes the drivers set (train.py:190-197, eval.py:64-69)
        self.temp = temp
        self.gate_layer = GlobalGate(branch_num=5)
        self.baseline = False
        self.ini_stage = False
        self.hard_gate = False
        self.save_weight_info = False
        self.weight_list = torch.Tensor()
        if encoder_rgb == "resnet34":
            self.flop = torch.tensor([0, 3.27, 7.27, 13.15, 16.02])
            self.depth_enc_flop = torch.tensor(DEPTH_ENC_FLOP_R34)
            self.total_flop = torch.tensor(TOTAL_FLOP_R34)
        else:
            self.depth_enc_flop = torch.tensor(DEPTH_ENC_FLOP_OTHER)
            self.total_flop = torch.tensor(TOTAL_FLOP_OTHER)
        # engine state
        self._cfg = dict(encoder=encoder_rgb, encoder_depth=encoder_depth, encoder_block=encoder_block,
                         fuse=fuse_depth_in_rgb_encoder, nr_decoder_blocks=tuple(nr_decoder_blocks),
                         num_classes=num_classes, upsampling=upsampling, context_module=context_module,
                         activation=activation, encoder_decoder_fusion=encoder_decoder_fusion)
        self._engine = None
        self._engine_key = None
        # arithmetic of the eval engine: "bf16" (stated tolerance 2e-2) or "f32x3" (fp32-grade: bf16 hi + lo operands,
        # three tensor-core products per MAC; logits within 1e-3 of the fp32 reference, ~3x the convolution work)
        self.engine_precision = "bf16"
        self._pending_weights: List[Tensor] = []
        self.use_cuda_graph = False          # opt-in: replay one captured graph per input shape
        # which captured-graph instance a forward replays: instances own their activation buffers, so forwards of
        # DIFFERENT instances launched on different streams overlap on the GPU (two batches in flight: the stem of
        # batch i+1 fills the SMs the latency-bound decoder tail of batch i leaves idle) -- see EvalPipeline
        self.graph_instance = 0
        # training arithmetic on CUDA: "fp32" = the reference's (library convs, our gate ops);
        # "bf16" = encoder/decoder convolutions forward + backward on the tcgen05 kernels (train_ops.py)
        self.train_precision = "fp32"
        self._graphs = {}

    # ------------------------------------------------------------------ reference API
    def freeze(self):                                       # :225-228
        for name, param in self.named_parameters():
            if "gate" not in name:
                param.requires_grad = False

    def start_weight(self):                                 # :230-232
        self.save_weight_info = True
        self.weight_list = torch.Tensor()
        self._pending_weights = []

    def _flush_weights(self):
        """Gate weights are accumulated on the device; the reference's per-forward
        ``weight.cpu()`` sync (:273-274) is deferred to here."""
        if self._pending_weights:
            new = torch.cat([w.detach().float().cpu() for w in self._pending_weights])
            self.weight_list = torch.cat((self.weight_list, new))
            self._pending_weights = []

    def end_weight(self, print_each=False, print_flop=False):   # :234-253
        self._flush_weights()
        self.save_weight_info = False
        if print_each:
            print(self.weight_list)
        stats = None
        if print_flop and self.weight_list.numel():
            cnt = np.array([(self.weight_list[:, i] == 1).sum().item() for i in range(5)], dtype=float)
            frac = torch.from_numpy(cnt / max(cnt.sum(), 1.0)).float()
            flop1 = (self.depth_enc_flop.cpu() * frac).sum()
            flop2 = (self.total_flop.cpu() * frac).sum()
            print(f"Depth Encoder Flop {flop1:.4f}G | Total Flop {flop2:.4f}G")
            stats = (cnt, float(flop1), float(flop2))
        self.weight_list = torch.Tensor()
        return stats

    # ------------------------------------------------------------------ engine plumbing
    def invalidate_engine(self):
        """Drop the packed weights; call after editing parameters in place while in eval mode.
        ``train()``, ``load_state_dict()`` and ``.to()/.cuda()`` do this automatically."""
        self._engine = None
        self._graphs = {}

    def train(self, mode=True):
        if mode:
            self.invalidate_engine()      # an optimizer is about to change the weights
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_engine()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_engine()
        out = super()._apply(fn, *args, **kwargs)
        for name in ("flop", "depth_enc_flop", "total_flop"):      # plain attributes in the reference (:217-223)
            if hasattr(self, name):
                setattr(self, name, fn(getattr(self, name)))
        return out

    def _state_version(self):
        # tripwire for in-place edits while in eval mode: the version counters of EVERY parameter and buffer
        # (~500 ints, negligible next to a forward).  Writes through ``.data`` do not bump them: call
        # invalidate_engine() after those.
        return tuple(p._version for p in self._version_probe)

    def engine(self, device=None):
        """The packed CUDA engine for the current weights (rebuilt when they change)."""
        from .engine import EngineConfig, FusionEngine
        device = device or next(self.parameters()).device
        if not hasattr(self, "_version_probe") or self._engine is None:
            self._version_probe = list(self.parameters()) + list(self.buffers())
        key = (str(device), self._state_version(), getattr(self, "engine_precision", "bf16"))
        if self._engine is None or self._engine_key != key:
            c = self._cfg
            cfg = EngineConfig(encoder=c["encoder"], encoder_depth=c["encoder_depth"],
                               encoder_block=c["encoder_block"], fuse=c["fuse"],
                               nr_decoder_blocks=c["nr_decoder_blocks"], num_classes=c["num_classes"],
                               upsampling=c["upsampling"], context_module=c["context_module"],
                               activation=c["activation"], precision=getattr(self, "engine_precision", "bf16"),
                               encoder_decoder_fusion=c["encoder_decoder_fusion"])
            self._engine = FusionEngine(self.state_dict(), cfg, device)
            self._engine_key = key
            self._graphs = {}
        return self._engine

    def _forward_engine(self, rgb, depth, labels_only=False, instance=None):
        eng = self.engine(rgb.device)
        instance = self.graph_instance if instance is None else instance
        modes = dict(temp=float(self.temp), hard_gate=bool(self.hard_gate), baseline=bool(self.baseline),
                     ini_stage=bool(self.ini_stage))
        if self.use_cuda_graph and not self.ini_stage:
            from .graph import GraphedForward
            key = (tuple(rgb.shape), tuple(sorted(modes.items())), labels_only, instance)
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = GraphedForward(eng, rgb, depth, modes, labels_only)
            return g(rgb, depth)
        if labels_only:
            b, _, h, w = rgb.shape
            labels = torch.empty(b, h, w, dtype=torch.uint8, device=rgb.device)
            _, weight = eng.forward(rgb, depth, labels=labels, want_logits=False, **modes)
            return labels, weight
        return eng.forward(rgb, depth, **modes)

    @torch.no_grad()
    def predict_labels(self, rgb, depth, out=None, return_weight=False, instance=None):
        """argmax_c(self(rgb, depth, True)) as uint8 [B,H,W] -- what eval.py:109-120 computes per batch --
        produced by the final upsampling kernel itself, so the 40-channel full-resolution logits are
        never written to memory.  Eval mode, CUDA tensors.  ``return_weight``: also the gate weights [B,5]
        (like ``forward(..., test=True, return_weight=True)``)."""
        if self.training or not rgb.is_cuda:
            raise RuntimeError("predict_labels runs the CUDA engine: call model.eval() and pass CUDA tensors")
        labels, weight = self._forward_engine(rgb, depth, labels_only=True, instance=instance)
        if self.save_weight_info:
            self._pending_weights.append(weight.detach().clone())
        if out is not None:
            out.copy_(labels)
            labels = out
        return (labels, weight) if return_weight else labels

    # ------------------------------------------------------------------ forward
    def _forward_eval_cuda(self, rgb, depth):
        """Eval forward on CUDA tensors: the engine.  A CONFIGURATION the engine does not implement (swish, appm,
        bilinear upsampling, mixed encoders, input not a multiple of 32, ...) is a NotImplementedError from the engine:
        such a model trained on the differentiable graph, so validation uses that graph too (one warning).  A missing
        library / wrong device is a DynmmError and always propagates -- there is no silent CPU or library fallback."""
        reason = getattr(self, "_engine_unsupported", None)
        if reason is None and (rgb.shape[2] % 32 or rgb.shape[3] % 32):
            if not getattr(self, "_warned_shape", False):
                self._warned_shape = True
                warnings.warn(f"dynmm_b200: input {tuple(rgb.shape[2:])} is not a multiple of 32; this eval forward runs "
                              f"the PyTorch graph instead of the CUDA engine")
            return self._forward_torch(rgb, depth)
        if reason is None:
            try:
                return self._forward_engine(rgb, depth)
            except NotImplementedError as e:
                reason = self._engine_unsupported = str(e) or "unsupported configuration"
                warnings.warn("dynmm_b200: the CUDA engine does not implement this configuration (" + reason +
                              "); eval forwards run the PyTorch graph instead")
        return self._forward_torch(rgb, depth)

    def forward(self, rgb, depth, test=False, return_weight=False):      # :255-322
        if rgb.is_cuda and not self.training:
            # (the reference's validate() wraps in no_grad; be safe when it is not)
            with torch.no_grad():
                out, weight = self._forward_eval_cuda(rgb, depth)
        else:
            out, weight = self._forward_torch(rgb, depth)
        if self.save_weight_info:
            self._pending_weights.append(weight.detach().clone())
        if test:
            return (out, weight) if return_weight else out
        flop = self.depth_enc_flop
        if flop.device != weight.device:
            flop = self.depth_enc_flop = flop.to(weight.device)
        loss = weight.mean(dim=0) * flop                     # FLOP regulariser, :314-315
        return out, loss.mean()

    def _forward_torch(self, rgb, depth):
        """Differentiable graph (training; also what runs for CPU tensors)."""
        se = self.fuse_depth_in_rgb_encoder != "add"
        if rgb.is_cuda and self.train_precision == "bf16":
            r = stem_channels_last(self.encoder_rgb, rgb)
            d = stem_channels_last(self.encoder_depth, depth)
        else:
            r = self.encoder_rgb.forward_first_conv(rgb)
            d = self.encoder_depth.forward_first_conv(depth)
        fuse = self.se_layer0(r, d) if se else r + d
        r = F.max_pool2d(fuse, 3, 2, 1)
        d = F.max_pool2d(d, 3, 2, 1)
        bs = r.shape[0]
        if self.baseline:
            weight = torch.zeros(bs, 5, device=r.device)
            weight[:, 4] = 1
        elif self.ini_stage:
            idx = torch.randint(0, 5, (bs,))
            weight = torch.zeros(bs, 5)
            weight[range(bs), idx] = 1
            weight = weight.to(r.device)
        else:
            weight = self.gate_layer(r, d, self.temp, self.hard_gate)
        g = torch.stack([1 - weight[:, 0], 1 - (weight[:, 0] + weight[:, 1]),
                         1 - (weight[:, 0] + weight[:, 1] + weight[:, 2]), weight[:, 4]])
        low = rgb.is_cuda and self.train_precision == "bf16"
        if self.train_precision not in ("fp32", "bf16"):
            raise ValueError("train_precision must be 'fp32' or 'bf16'")
        if low:
            # stem + gate stay fp32 (decisions must match the reference); everything after runs bf16 NHWC
            r = r.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
            d = d.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        fused = []
        for s in range(4):
            r = getattr(self.encoder_rgb, f"layer{s + 1}")(r if s == 0 else fuse)
            d = getattr(self.encoder_depth, f"layer{s + 1}")(d)
            if se:
                # w*rgb + (1-w)*se(rgb,depth)  with w = 1 - g_s
                b1 = getattr(self, f"se_layer{s + 1}")(r, d)
                gs = g[s].view(-1, 1, 1, 1).to(r.dtype)
                fuse = (1 - gs) * r + gs * b1
            elif low:
                fuse = r + g[s].view(-1, 1, 1, 1).to(r.dtype) * d
            else:
                fuse = gated_blend(r, d, g[s])              # rgb + g_s * depth
            fused.append(fuse)
        skips = [getattr(self, f"skip_layer{i + 1}")(fused[i]) for i in range(3)]
        out = self.decoder([self.context_module(fused[3]), skips[2], skips[1], skips[0]])
        if low:
            out = tuple(o.float() for o in out) if isinstance(out, tuple) else out.float()
        return out, weight
