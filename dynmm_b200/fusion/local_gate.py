"""Local-gate dynamic ESANet: ``SkipESANet`` with one two-way gate per fusion site.

Drop-in surface of ``FusionDynMM/src/models/model_skip_mod.py:20-324`` (``SkipESANet``),
``rgb_depth_fusion.py:29-65`` (``SqueezeAndExciteReweigh``) and ``model_utils.py:54-70``
(``SqueezeAndExcitationWeight``): same constructor arguments, ``state_dict`` keys, mode attributes
(``hard_gate, ini_stage, random_policy, save_weight_info, weight_list, block_rule, temp``) and return convention
(``forward(rgb, depth, test=False) -> logits`` in eval mode, the 4-scale tuple in training mode).

What the reference computes per fusion site s = 0..3 (gate s is evaluated on the features ENTERING stage s+1 and
decides the blend AFTER stage s+1, model_skip_mod.py:240-311):

    x      = cat(rgb_s, depth_s)                                   [B, 2C, h, w]
    score  = sigmoid(mean_{c,h,w}(x * SE(x)))                       (SqueezeAndExcitationWeight + Sigmoid)
    w_s    = gumbel_softmax(stack(score, 1 - score) / temp, hard)   (real Gumbel noise; hard=True when test=True)
    w_s[1] = w_s[1] * w_{s-1}[1]   (chained unless ini_stage),  w_s[0] = 1 - w_s[1]
    fuse   = w_s[0] * rgb_{s+1} + w_s[1] * (rgb_{s+1} + depth_{s+1})      if block_rule[s] == 2, else rgb / rgb+depth

Notes that matter for a re-implementation:
  * ``mean(x * SE(x))`` only needs the global average pool of x:  (1/2C) sum_c SE(gap)_c * gap_c  -- the concatenated
    tensor and the scaled copy are never built here;
  * the blend is ``rgb + w_s[1] * depth`` (w_s[0] + w_s[1] == 1): on CUDA fp32 tensors it runs on the same
    ``gated_add`` kernel (with hand-written backward) as the global-gate model;
  * the Gumbel-softmax is ``DiffSoftmax(logits + g, tau=1, hard)`` with ``g = -log(Exp(1))`` drawn exactly like
    ``F.gumbel_softmax`` draws it (same generator, same order, same shape), so on the same device and seed the decisions
    equal the reference's; the softmax / first-max one-hot / straight-through part is our ``diff_softmax`` op;
  * ``se_layer0..4`` are constructed for ``fuse_depth_in_rgb_encoder='SE-add'`` (they are in the state_dict) but the
    reference's forward never calls them (model_skip_mod.py:241-311 always adds) -- kept that way;
  * ``SqueezeAndExcitationWeight.linear`` is a parameter the reference never uses (model_utils.py:64); kept for the keys.

Two execution paths, same surface:
  * the differentiable PyTorch graph (training mode, CPU tensors, ``use_engine = False``): parity with the
    reference-generated vectors (tests/golden/local_gate_*.npz); convolutions optionally on the tcgen05 kernels with
    ``train_precision='bf16'`` exactly like ``SkipGateESANet``'s training graph (modules.Conv2d);
  * eval mode on CUDA tensors: ``FusionEngine.forward_local`` (bf16 NHWC kernels of the global-gate engine) with REAL
    per-stage skipping.  The local gates' decisions are not known before stage 1 (unlike the global gate's), but site s
    decides before stage s + 1 runs and chained hard gates (w_{s-1}[1] == 0 closes all later sites,
    rgb_depth_fusion.py:60-63) make the set of samples that can still use depth features shrink monotonically: every
    stage re-plans its slot order on the device from the gate weights, the depth encoder's convolutions run on the
    ``count`` prefix only.  Configurations the engine does not implement (bilinear up-sampling, swish, ...) run the
    PyTorch graph, with one warning.
"""
from __future__ import annotations

import warnings
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .autograd_ops import diff_softmax, gated_blend
from .modules import (Conv2d, ConvBNAct, Decoder, SqueezeAndExciteFusionAdd, _build_encoder, _make_activation,
                      get_context_module)

Tensor = torch.Tensor


class SqueezeAndExcitationWeight(nn.Module):
    """model_utils.py:54-70 (keys: fc.0, fc.2, linear).  Returns mean_{c,h,w}(x * SE(x)) per sample."""

    def __init__(self, channel, reduction=16, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.fc = nn.Sequential(Conv2d(channel, channel // reduction, 1), activation,
                                Conv2d(channel // reduction, channel, 1), nn.Sigmoid())
        self.linear = nn.Linear(channel, 2)          # unused by the reference's forward; present in its state_dict

    def forward_pooled(self, gap: Tensor) -> Tensor:
        """gap: [B, C, 1, 1] global average pool of x (fp32)."""
        return (gap * self.fc(gap)).mean(dim=(1, 2, 3))

    def forward(self, x):
        return self.forward_pooled(F.adaptive_avg_pool2d(x.float(), 1))


def gumbel_softmax(logits: Tensor, hard: bool) -> Tensor:
    """``F.gumbel_softmax(logits, tau=1, hard=hard)`` with the noise drawn like torch draws it and the
    softmax / arg-max / straight-through part on ``diff_softmax`` (custom CUDA forward + backward on the GPU)."""
    gumbels = -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()
    return diff_softmax(logits + gumbels, 1.0, hard, -1)


class SqueezeAndExciteReweigh(nn.Module):
    """rgb_depth_fusion.py:29-65: two-way gate of one fusion site -> [B, 2, 1, 1]."""

    def __init__(self, temp, channels_in, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.temp = temp
        self.se = SqueezeAndExcitationWeight(channels_in * 2, activation=activation)
        self.act = nn.Sigmoid()

    def forward(self, rgb, depth, hard=False, prev_weight=None, random=False, test=False):
        if random:
            bs = rgb.shape[0]
            b0 = torch.randint(0, 2, (bs,))                     # global CPU generator (:38)
            w_norm = torch.stack([b0, 1 - b0], dim=1).to(rgb.device)
        else:
            # global average pool of cat(rgb, depth) without building the concatenation
            gap = torch.cat([F.adaptive_avg_pool2d(rgb.float(), 1), F.adaptive_avg_pool2d(depth.float(), 1)], dim=1)
            w = self.act(self.se.forward_pooled(gap))
            w = torch.stack([w, 1 - w], dim=1)
            w_norm = gumbel_softmax(w / self.temp, hard=True if test else hard)
        if prev_weight is not None:
            b1 = w_norm[:, 1] * prev_weight
            w_norm = torch.stack([1 - b1, b1], dim=1)
        return w_norm.view(-1, 2, 1, 1)


class SkipESANet(nn.Module):
    """Local-gate dynamic ESANet (model_skip_mod.py:20-324)."""

    def __init__(self, height=480, width=640, num_classes=37, encoder_rgb="resnet18", encoder_depth="resnet18",
                 encoder_block="BasicBlock", channels_decoder=None, pretrained_on_imagenet=False,
                 pretrained_dir="./trained_models/imagenet", activation="relu", encoder_decoder_fusion="add",
                 context_module="ppm", nr_decoder_blocks=None, fuse_depth_in_rgb_encoder="SE-add",
                 upsampling="bilinear", temp=1, block_rule=None):
        super().__init__()
        channels_decoder = [128, 128, 128] if channels_decoder is None else list(channels_decoder)
        nr_decoder_blocks = [1, 1, 1] if nr_decoder_blocks is None else list(nr_decoder_blocks)
        self._cfg = dict(encoder=encoder_rgb, encoder_depth=encoder_depth, encoder_block=encoder_block,
                         nr_decoder_blocks=tuple(nr_decoder_blocks), num_classes=num_classes, upsampling=upsampling,
                         context_module=context_module, activation=activation,
                         encoder_decoder_fusion=encoder_decoder_fusion)
        self.fuse_depth_in_rgb_encoder = fuse_depth_in_rgb_encoder
        # 0: rgb only, 1: rgb + depth, 2: dynamic (:38, :49)
        self.block_rule = block_rule if block_rule else [1, 1, 1, 1]
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
        if fuse_depth_in_rgb_encoder == "SE-add":                # built, never called by forward (:124-139, :241-311)
            for i, c in enumerate((64,) + stage_ch):
                setattr(self, f"se_layer{i}", SqueezeAndExciteFusionAdd(c, activation=self.activation))
        self.temp = temp
        for i, c in enumerate((64,) + stage_ch[:3]):             # :141-146
            setattr(self, f"gate_layer{i}", SqueezeAndExciteReweigh(self.temp, c, activation=self.activation))
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
        self.hard_gate = False
        self.ini_stage = False
        self.random_policy = False
        self.save_weight_info = False
        self.weight_list = [torch.Tensor() for _ in range(4)]
        self._pending: List[List[Tensor]] = [[] for _ in range(4)]
        self.train_precision = "fp32"        # "bf16": stage convolutions on the tcgen05 kernels (modules.Conv2d)
        self.use_engine = True               # eval mode + CUDA tensors: FusionEngine.forward_local (real skipping)
        self.engine_precision = "bf16"       # "f32x3": fp32-grade engine arithmetic (logits within 1e-3 of fp32)
        self.use_cuda_graph = False          # opt-in: replay one captured graph per input shape / mode (static outputs)
        self._graphs = {}
        self._engine = None
        self.last_counts = None              # device int32 [1] x 4: depth samples each stage processed (engine path)

    # ------------------------------------------------------------------ reference API
    def freeze(self):                                             # :215-218
        for name, param in self.named_parameters():
            if "gate" not in name:
                param.requires_grad = False

    def start_weight(self):                                       # :220-222
        self.save_weight_info = True
        self.weight_list = [torch.Tensor() for _ in range(4)]
        self._pending = [[] for _ in range(4)]

    def _flush_weights(self):
        """Gate weights are kept on their device per forward; the reference's per-forward ``.cpu()`` syncs
        (:247-248, :265-266, ...) happen once, here."""
        for i in range(4):
            if self._pending[i]:
                new = torch.cat([w.detach().float().cpu() for w in self._pending[i]])
                self.weight_list[i] = torch.cat((self.weight_list[i], new))
                self._pending[i] = []

    def end_weight(self, print_each=False, thre=None):           # :224-243
        self._flush_weights()
        self.save_weight_info = False
        avg = []
        for i in range(4):
            if self.block_rule[i] != 2:
                continue
            if thre:
                print("-" * 40, "layer ", i, "-" * 40)
                cnt1 = (self.weight_list[i][:, 0] < thre).sum()
                cnt2 = (self.weight_list[i][:, 1] < thre).sum()
                print(f"Skip {cnt1} branch 1 | {cnt2} branch 2")
            weight_mean = torch.mean(self.weight_list[i], axis=0)
            if print_each:
                print(self.weight_list[i])
                print(weight_mean)
            avg.append(weight_mean)
        self.weight_list = [torch.Tensor() for _ in range(4)]
        return avg

    # ------------------------------------------------------------------ forward
    def _blend(self, rule: int, weight: Tensor, rgb: Tensor, depth: Tensor) -> Tensor:
        if rule == 0:
            return rgb
        if rule == 1:
            return rgb + depth
        if rgb.is_cuda and rgb.dtype == torch.float32:
            # w0*rgb + w1*(rgb+depth) with w0 + w1 == 1  ->  rgb + w1*depth on the gated_add kernel
            return gated_blend(rgb, depth, weight[:, 1, 0, 0])
        w = weight.to(rgb.dtype)
        return w[:, 0:1] * rgb + w[:, 1:2] * (rgb + depth)       # the reference's statement (:258, :276, :294, :309)

    def engine(self, device=None):
        """The packed CUDA engine for the current weights (rebuilt when a parameter or buffer changes)."""
        from .engine import EngineConfig, FusionEngine
        device = device or next(self.parameters()).device
        if self._engine is None:
            self._version_probe = list(self.parameters()) + list(self.buffers())
        key = (str(device), tuple(p._version for p in self._version_probe), getattr(self, "engine_precision", "bf16"))
        if self._engine is None or self._engine_key != key:
            c = self._cfg
            # the reference's forward always blends by addition, whatever fuse_depth_in_rgb_encoder built (:241-311)
            cfg = EngineConfig(encoder=c["encoder"], encoder_depth=c["encoder_depth"],
                               encoder_block=c["encoder_block"], fuse="add",
                               nr_decoder_blocks=c["nr_decoder_blocks"], num_classes=c["num_classes"],
                               upsampling=c["upsampling"], context_module=c["context_module"],
                               activation=c["activation"], gate="local",
                               encoder_decoder_fusion=c["encoder_decoder_fusion"],
                               precision=getattr(self, "engine_precision", "bf16"))
            self._engine = FusionEngine(self.state_dict(), cfg, device)
            self._engine_key = key
            self._graphs = {}
        return self._engine

    def invalidate_engine(self):
        self._engine = None
        self._graphs = {}

    def train(self, mode: bool = True):
        if mode:
            self._engine = None            # an optimizer is about to change the weights
        return super().train(mode)

    def load_state_dict(self, *args, **kw):
        self._engine = None
        return super().load_state_dict(*args, **kw)

    def _forward_engine(self, rgb, depth, test):
        eng = self.engine(rgb.device)
        modes = dict(block_rule=tuple(self.block_rule), temp=float(self.gate_layer0.temp),
                     hard=True if test else bool(self.hard_gate), random_policy=bool(self.random_policy),
                     ini_stage=bool(self.ini_stage))
        if self.use_cuda_graph and not self.random_policy:
            from .graph import GraphedLocalForward
            key = (tuple(rgb.shape), tuple(sorted(modes.items())))
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = GraphedLocalForward(eng, rgb, depth, modes)
            out, weights, counts = g(rgb, depth)
        else:
            out, weights, counts = eng.forward_local(rgb, depth, **modes)
        self.last_counts = counts
        if self.save_weight_info:
            for i in range(4):
                self._pending[i].append(weights[i].detach())
        return out

    def forward(self, rgb, depth, test=False):                   # :246-322
        if self.train_precision not in ("fp32", "bf16"):
            raise ValueError("train_precision must be 'fp32' or 'bf16'")
        if self.use_engine and not self.training and rgb.is_cuda and not torch.is_grad_enabled() and \
                getattr(self, "_engine_unsupported", None) is None and rgb.shape[2] % 32 == 0 and rgb.shape[3] % 32 == 0:
            # A configuration the engine does not implement is a NotImplementedError: such a model runs the PyTorch
            # graph below (one warning).  A missing library / wrong device is a DynmmError and propagates.
            try:
                return self._forward_engine(rgb, depth, test)
            except NotImplementedError as e:
                self._engine_unsupported = str(e) or "unsupported configuration"
                warnings.warn("dynmm_b200: the CUDA engine does not implement this configuration (" +
                              self._engine_unsupported + "); eval forwards run the PyTorch graph instead")
        low = rgb.is_cuda and self.train_precision == "bf16"
        gate_args = dict(hard=self.hard_gate, random=self.random_policy, test=test)
        if low:
            from .modules import stem_channels_last
            rgb = stem_channels_last(self.encoder_rgb, rgb)
            depth = stem_channels_last(self.encoder_depth, depth)
        else:
            rgb = self.encoder_rgb.forward_first_conv(rgb)
            depth = self.encoder_depth.forward_first_conv(depth)
        fuse = rgb + depth
        weights: List[Tensor] = [self.gate_layer0(rgb, depth, **gate_args)]
        rgb = F.max_pool2d(fuse, kernel_size=3, stride=2, padding=1)
        depth = F.max_pool2d(depth, kernel_size=3, stride=2, padding=1)
        if low:      # the stems stay fp32 like in SkipGateESANet; everything after runs bf16 NHWC
            rgb = rgb.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
            depth = depth.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        prev_weight: Optional[Tensor] = None
        fused = []
        fuse = rgb
        for s in range(4):
            rgb = getattr(self.encoder_rgb, f"forward_layer{s + 1}")(fuse)
            depth = getattr(self.encoder_depth, f"forward_layer{s + 1}")(depth)
            fuse = self._blend(self.block_rule[s], weights[s], rgb, depth)
            if self.block_rule[s] not in (0, 1):
                prev_weight = weights[s][:, 1, 0, 0] if not self.ini_stage else None
            fused.append(fuse)
            if s < 3:
                # gate s+1 looks at the stage outputs BEFORE the blend (:260, :278, :296)
                weights.append(getattr(self, f"gate_layer{s + 1}")(rgb, depth, prev_weight=prev_weight, **gate_args))
        if self.save_weight_info:
            for i in range(4):
                self._pending[i].append(weights[i][:, :, 0, 0].detach())
        skips = [getattr(self, f"skip_layer{i + 1}")(fused[i]) for i in range(3)]
        out = self.decoder([self.context_module(fused[3]), skips[2], skips[1], skips[0]])
        if low:
            out = tuple(o.float() for o in out) if isinstance(out, tuple) else out.float()
        return out
