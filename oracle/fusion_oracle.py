"""CPU oracle (fp32, plain ``torch.nn.functional``) for fusion-level DynMM.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This is a functional
restatement of the reference's gated RGB-D forward, driven by a flat
``state_dict`` (the reference's key names), so the same tensors can be loaded
into the reference module, into this oracle and into the CUDA engine.

Pinned against the reference itself: ``oracle/make_golden.py`` imports
``/root/reference/FusionDynMM`` in the build container, loads the seeded
``state_dict`` from :func:`make_state_dict` into the reference
``SkipGateESANet`` (strict) and stores its outputs in ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors.

Reference citations are relative to ``/root/reference/FusionDynMM/src/models``.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# FLOP tables quoted by the reference (model_skip_mod_globalgate.py:217-223).
DEPTH_ENC_FLOP_R34 = (0.2506752, 3.1113216, 6.9470208, 12.66432, 15.538944)
TOTAL_FLOP_R34 = (22.37101509, 25.23166149, 29.06736069, 34.78465989, 37.65928389)
DEPTH_ENC_FLOP_OTHER = (0.2506752, 4.39420573, 10.72382115, 19.71582947, 24.679084)
TOTAL_FLOP_OTHER = (32.5854654, 36.728995928, 43.058611352, 52.050619672, 57.0138742)

_STAGE_BLOCKS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}
_STAGE_PLANES = (64, 128, 256, 512)


@dataclasses.dataclass(frozen=True)
class FusionConfig:
    """Constructor arguments of the reference model that change the arithmetic
    (model_skip_mod_globalgate.py:34-51)."""
    height: int = 480
    width: int = 640
    num_classes: int = 40
    encoder: str = "resnet34"
    encoder_block: str = "NonBottleneck1D"
    channels_decoder: Sequence[int] = (128, 128, 128)
    nr_decoder_blocks: Sequence[int] = (3, 3, 3)
    fuse_depth_in_rgb_encoder: str = "add"          # 'add' | 'SE-add'
    context_module: str = "ppm"
    upsampling: str = "learned-3x3-zeropad"
    activation: str = "relu"

    @property
    def block(self) -> str:
        # resnet.py:443-446: ResNet50 ignores encoder_block and uses Bottleneck.
        return "Bottleneck" if self.encoder == "resnet50" else self.encoder_block

    @property
    def expansion(self) -> int:
        return 4 if self.block == "Bottleneck" else 1

    @property
    def stage_channels(self):
        return tuple(p * self.expansion for p in _STAGE_PLANES)

    @property
    def stage_blocks(self):
        return _STAGE_BLOCKS[self.encoder]


# --------------------------------------------------------------------------
# seeded weights (identical key names / shapes as the reference state_dict)
# --------------------------------------------------------------------------

def _param_specs(cfg: FusionConfig):
    """Yield (key, shape, kind) for every entry of the reference state_dict.

    kind in {conv, bias, bn_w, bn_b, bn_m, bn_v, bn_n, upw}.
    """
    out = []

    def conv(key, cout, cin, kh, kw, bias):
        out.append((key + ".weight", (cout, cin, kh, kw), "conv"))
        if bias:
            out.append((key + ".bias", (cout,), "bias"))

    def bn(key, c):
        out.append((key + ".weight", (c,), "bn_w"))
        out.append((key + ".bias", (c,), "bn_b"))
        out.append((key + ".running_mean", (c,), "bn_m"))
        out.append((key + ".running_var", (c,), "bn_v"))
        out.append((key + ".num_batches_tracked", (), "bn_n"))

    def nbt1d(key, cin, c):
        conv(key + ".conv3x1_1", c, cin, 3, 1, True)
        conv(key + ".conv1x3_1", c, c, 1, 3, True)
        bn(key + ".bn1", c)
        conv(key + ".conv3x1_2", c, c, 3, 1, True)
        conv(key + ".conv1x3_2", c, c, 1, 3, True)
        bn(key + ".bn2", c)

    def conv_bn_act(key, cin, cout, k):
        conv(key + ".conv", cout, cin, k, k, False)
        bn(key + ".bn", cout)

    def encoder(key, cin0):
        conv(key + ".conv1", 64, cin0, 7, 7, False)
        bn(key + ".bn1", 64)
        inplanes = 64
        for s, (planes, nblk) in enumerate(zip(_STAGE_PLANES, cfg.stage_blocks)):
            stride = 1 if s == 0 else 2
            for b in range(nblk):
                bk = f"{key}.layer{s + 1}.{b}"
                cin = inplanes if b == 0 else planes * cfg.expansion
                if cfg.block == "NonBottleneck1D":
                    nbt1d(bk, cin, planes)
                elif cfg.block == "BasicBlock":
                    conv(bk + ".conv1", planes, cin, 3, 3, False)
                    bn(bk + ".bn1", planes)
                    conv(bk + ".conv2", planes, planes, 3, 3, False)
                    bn(bk + ".bn2", planes)
                else:  # Bottleneck, resnet.py:150-171
                    conv(bk + ".conv1", planes, cin, 1, 1, False)
                    bn(bk + ".bn1", planes)
                    conv(bk + ".conv2", planes, planes, 3, 3, False)
                    bn(bk + ".bn2", planes)
                    conv(bk + ".conv3", planes * 4, planes, 1, 1, False)
                    bn(bk + ".bn3", planes * 4)
                if b == 0 and (stride != 1 or inplanes != planes * cfg.expansion):
                    conv(bk + ".downsample.0", planes * cfg.expansion, inplanes, 1, 1, False)
                    bn(bk + ".downsample.1", planes * cfg.expansion)
            inplanes = planes * cfg.expansion

    encoder("encoder_rgb", 3)
    encoder("encoder_depth", 1)
    ch = cfg.stage_channels
    if cfg.fuse_depth_in_rgb_encoder == "SE-add":
        for i, c in enumerate((64,) + ch):
            for m in ("se_rgb", "se_depth"):
                conv(f"se_layer{i}.{m}.fc.0", c // 16, c, 1, 1, True)
                conv(f"se_layer{i}.{m}.fc.2", c, c // 16, 1, 1, True)
    dec = list(cfg.channels_decoder)
    for i, (c_enc, c_dec) in enumerate(zip(ch[:3], (dec[2], dec[1], dec[0]))):
        if c_enc != c_dec:
            conv_bn_act(f"skip_layer{i + 1}.0", c_enc, c_dec, 1)
    # context module: ppm with bins (1, 5) (context_modules.py:28-38)
    cin = ch[3]
    red = cin // 2
    for i in range(2):
        conv_bn_act(f"context_module.features.{i}.1", cin, red, 1)
    conv_bn_act("context_module.final_conv", cin + 2 * red, dec[0], 1)
    # decoder (model.py:244-357)
    c_prev = dec[0]
    for i in range(3):
        dk = f"decoder.decoder_module_{i + 1}"
        conv_bn_act(dk + ".conv3x3", c_prev, dec[i], 3)
        for b in range(cfg.nr_decoder_blocks[i]):
            nbt1d(f"{dk}.decoder_blocks.{b}", dec[i], dec[i])
        out.append((dk + ".upsample.conv.weight", (dec[i], 1, 3, 3), "upw"))
        out.append((dk + ".upsample.conv.bias", (dec[i],), "bias"))
        conv(dk + ".side_output", cfg.num_classes, dec[i], 1, 1, True)
        c_prev = dec[i]
    conv("decoder.conv_out", cfg.num_classes, dec[2], 3, 3, True)
    for u in ("upsample1", "upsample2"):
        out.append((f"decoder.{u}.conv.weight", (cfg.num_classes, 1, 3, 3), "upw"))
        out.append((f"decoder.{u}.conv.bias", (cfg.num_classes,), "bias"))
    # gate (model_skip_mod_globalgate.py:375-386)
    conv("gate_layer.conv.0", 8, 128, 5, 5, True)
    bn("gate_layer.conv.1", 8)
    conv("gate_layer.conv.3", 8, 8, 5, 5, True)
    bn("gate_layer.conv.4", 8)
    conv("gate_layer.fc", 5, 8, 1, 1, False)
    return out


def make_state_dict(cfg: FusionConfig, seed: int = 0, gate_scale: float = 1.0) -> SD:
    """Deterministic, non-trivial weights: He-style convs, random BN affine AND
    running statistics (so eval-mode BN is not an identity), random biases.
    ``gate_scale`` widens the gate fc weights so an untrained gate still spreads
    samples over several branches."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for key, shape, kind in _param_specs(cfg):
        if kind == "conv":
            cout, cin, kh, kw = shape
            std = math.sqrt(2.0 / (cin * kh * kw))
            t = torch.randn(shape, generator=g) * std
            if key == "gate_layer.fc.weight":
                t = t * gate_scale
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            t = 0.6 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_m":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_v":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "bn_n":
            t = torch.tensor(0, dtype=torch.long)
        elif kind == "upw":
            # the reference initialises to a fixed bilinear-like stencil
            # (model.py:385-391); perturb so the conv is exercised generally.
            w = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]])
            t = w.expand(shape).clone() * (1.0 + 0.2 * torch.randn(shape, generator=g))
        else:
            raise AssertionError(kind)
        sd[key] = t
    return sd


# --------------------------------------------------------------------------
# arithmetic
# --------------------------------------------------------------------------

def diff_softmax(logits: Tensor, tau: float = 1.0, hard: bool = False, dim: int = -1) -> Tensor:
    """model_skip_mod_globalgate.py:20-30 (same in imdb_dyn.py:16-26,
    affect_dyn.py:18-28): tempered softmax; hard = one-hot of the FIRST max of
    y_soft with a straight-through gradient."""
    y_soft = torch.softmax(logits / tau, dim)
    if not hard:
        return y_soft
    idx = y_soft.max(dim, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(dim, idx, 1.0)
    return y_hard - y_soft.detach() + y_soft


def _act(x: Tensor, name: str) -> Tensor:
    if name == "relu":
        return F.relu(x)
    if name in ("swish", "silu"):
        return x * torch.sigmoid(x)            # model_utils.py:105-106
    if name == "hswish":
        return x * F.relu6(x + 3.0) / 6.0      # model_utils.py:114-115
    raise NotImplementedError(name)


class _Ctx:
    """BatchNorm mode + running-stat bookkeeping for one oracle call."""

    def __init__(self, sd: SD, training: bool, act: str):
        self.sd, self.training, self.act = sd, training, act

    def bn(self, x: Tensor, key: str, eps: float = 1e-5) -> Tensor:
        sd = self.sd
        if self.training:
            return F.batch_norm(x, None, None, sd[key + ".weight"], sd[key + ".bias"], True, 0.1, eps)
        return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                            sd[key + ".weight"], sd[key + ".bias"], False, 0.1, eps)

    def conv(self, x: Tensor, key: str, stride=1, padding=0, groups=1) -> Tensor:
        return F.conv2d(x, self.sd[key + ".weight"], self.sd.get(key + ".bias"), stride, padding, 1, groups)

    def a(self, x: Tensor) -> Tensor:
        return _act(x, self.act)


def nbt1d_block(c: _Ctx, key: str, x: Tensor, stride: int = 1) -> Tensor:
    """resnet.py:124-147 -- factorised 3x1/1x3 residual block, BN eps 1e-3,
    stride applied as (s,1) on the first 3x1 and (1,s) on the first 1x3."""
    y = c.conv(x, key + ".conv3x1_1", (stride, 1), (1, 0))
    y = c.a(y)
    y = c.conv(y, key + ".conv1x3_1", (1, stride), (0, 1))
    y = c.a(c.bn(y, key + ".bn1", 1e-3))
    y = c.a(c.conv(y, key + ".conv3x1_2", 1, (1, 0)))
    y = c.bn(c.conv(y, key + ".conv1x3_2", 1, (0, 1)), key + ".bn2", 1e-3)
    if key + ".downsample.0.weight" in c.sd:
        idn = c.bn(c.conv(x, key + ".downsample.0", stride), key + ".downsample.1")
    else:
        idn = x
    return c.a(y + idn)


def basic_block(c: _Ctx, key: str, x: Tensor, stride: int = 1) -> Tensor:
    """resnet.py:66-84."""
    y = c.a(c.bn(c.conv(x, key + ".conv1", stride, 1), key + ".bn1"))
    y = c.bn(c.conv(y, key + ".conv2", 1, 1), key + ".bn2")
    if key + ".downsample.0.weight" in c.sd:
        idn = c.bn(c.conv(x, key + ".downsample.0", stride), key + ".downsample.1")
    else:
        idn = x
    return c.a(y + idn)


def bottleneck_block(c: _Ctx, key: str, x: Tensor, stride: int = 1) -> Tensor:
    """resnet.py:173-192."""
    y = c.a(c.bn(c.conv(x, key + ".conv1"), key + ".bn1"))
    y = c.a(c.bn(c.conv(y, key + ".conv2", stride, 1), key + ".bn2"))
    y = c.bn(c.conv(y, key + ".conv3"), key + ".bn3")
    if key + ".downsample.0.weight" in c.sd:
        idn = c.bn(c.conv(x, key + ".downsample.0", stride), key + ".downsample.1")
    else:
        idn = x
    return c.a(y + idn)


_BLOCKS = {"NonBottleneck1D": nbt1d_block, "BasicBlock": basic_block, "Bottleneck": bottleneck_block}


def encoder_first_conv(c: _Ctx, key: str, x: Tensor) -> Tensor:
    """resnet.py:352-358: conv7x7 s2 p3 (no bias) + BN + act."""
    return c.a(c.bn(c.conv(x, key + ".conv1", 2, 3), key + ".bn1"))


def encoder_layer(c: _Ctx, cfg: FusionConfig, key: str, stage: int, x: Tensor) -> Tensor:
    """resnet.py:360-379 (forward_layer{stage}); stage is 1-based."""
    blk = _BLOCKS[cfg.block]
    for b in range(cfg.stage_blocks[stage - 1]):
        x = blk(c, f"{key}.layer{stage}.{b}", x, 2 if (b == 0 and stage > 1) else 1)
    return x


def squeeze_excite(c: _Ctx, key: str, x: Tensor) -> Tensor:
    """model_utils.py:47-51."""
    w = F.adaptive_avg_pool2d(x, 1)
    w = c.a(c.conv(w, key + ".fc.0"))
    w = torch.sigmoid(c.conv(w, key + ".fc.2"))
    return x * w


def se_fusion_add(c: _Ctx, key: str, rgb: Tensor, depth: Tensor) -> Tensor:
    """rgb_depth_fusion.py:22-26."""
    return squeeze_excite(c, key + ".se_rgb", rgb) + squeeze_excite(c, key + ".se_depth", depth)


def conv_bn_act(c: _Ctx, key: str, x: Tensor, k: int) -> Tensor:
    """model_utils.py:11-23."""
    return c.a(c.bn(c.conv(x, key + ".conv", 1, k // 2), key + ".bn"))


def global_gate_logits(c: _Ctx, rgb: Tensor, depth: Tensor) -> Tensor:
    """model_skip_mod_globalgate.py:388-392, up to (not including) DiffSoftmax.
    Returns [B,5]."""
    x = torch.cat([rgb, depth], 1)
    y = torch.tanh(c.bn(c.conv(x, "gate_layer.conv.0", 2), "gate_layer.conv.1"))
    y = torch.tanh(c.bn(c.conv(y, "gate_layer.conv.3", 2), "gate_layer.conv.4"))
    y = F.adaptive_avg_pool2d(y, 1)
    y = c.conv(y, "gate_layer.fc")
    return y.flatten(1)


def context_ppm(c: _Ctx, x: Tensor, upsampling: str) -> Tensor:
    """context_modules.py:69-87 with bins (1,5).  'learned-3x3*' upsampling
    degrades to nearest for the context module
    (model_skip_mod_globalgate.py:180-187)."""
    mode = "nearest" if "learned-3x3" in upsampling else upsampling
    h, w = x.shape[2:]
    outs = [x]
    for i, b in enumerate((1, 5)):
        y = F.adaptive_avg_pool2d(x, b)
        y = conv_bn_act(c, f"context_module.features.{i}.1", y, 1)
        if mode == "nearest":
            outs.append(F.interpolate(y, (h, w), mode="nearest"))
        else:
            outs.append(F.interpolate(y, (h, w), mode="bilinear", align_corners=False))
    return conv_bn_act(c, "context_module.final_conv", torch.cat(outs, 1), 1)


def upsample2x(c: _Ctx, key: str, x: Tensor, mode: str) -> Tensor:
    """model.py:403-410: nearest x2 then depthwise 3x3 (zero pad) for
    'learned-3x3-zeropad'; replication pad for 'learned-3x3'."""
    size = (x.shape[2] * 2, x.shape[3] * 2)
    if "learned-3x3" in mode:
        x = F.interpolate(x, size, mode="nearest")
        ch = x.shape[1]
        if mode == "learned-3x3":
            x = F.pad(x, (1, 1, 1, 1), mode="replicate")
            return c.conv(x, key + ".conv", 1, 0, ch)
        return c.conv(x, key + ".conv", 1, 1, ch)
    if mode == "bilinear":
        return F.interpolate(x, size, mode="bilinear", align_corners=False)
    return F.interpolate(x, size, mode=mode)


def decoder(c: _Ctx, cfg: FusionConfig, enc_outs: List[Tensor]):
    """model.py:295-308 and DecoderModule.forward :341-357."""
    out, s16, s8, s4 = enc_outs
    sides = []
    for i, skip in enumerate((s16, s8, s4)):
        dk = f"decoder.decoder_module_{i + 1}"
        out = conv_bn_act(c, dk + ".conv3x3", out, 3)
        for b in range(cfg.nr_decoder_blocks[i]):
            out = nbt1d_block(c, f"{dk}.decoder_blocks.{b}", out)
        sides.append(c.conv(out, dk + ".side_output") if c.training else None)
        out = upsample2x(c, dk + ".upsample", out, cfg.upsampling)
        out = out + skip
    out = c.conv(out, "decoder.conv_out", 1, 1)
    out = upsample2x(c, "decoder.upsample1", out, cfg.upsampling)
    out = upsample2x(c, "decoder.upsample2", out, cfg.upsampling)
    if c.training:
        return out, sides[2], sides[1], sides[0]
    return out


def stage_gates(weight: Tensor) -> Tensor:
    """Per-stage depth mixing coefficient g_s with fuse_s = rgb_s + g_s*depth_s
    ('add' fusion).  model_skip_mod_globalgate.py:282,291,300,309: stages 1-3
    blend with w = sum_{k<s} weight[:,k] on the RGB-only branch, stage 4 with
    weight[:,4] on the fused branch.  Returns [B,4] of (1-w_1, 1-w_2, 1-w_3,
    weight[:,4])."""
    w1 = weight[:, 0]
    w2 = weight[:, 0] + weight[:, 1]
    w3 = weight[:, 0] + weight[:, 1] + weight[:, 2]
    return torch.stack([1 - w1, 1 - w2, 1 - w3, weight[:, 4]], 1)


def forward(sd: SD, cfg: FusionConfig, rgb: Tensor, depth: Tensor, *, temp: float = 1.0,
            hard_gate: bool = False, baseline: bool = False, ini_stage: bool = False,
            training: bool = False, weight: Optional[Tensor] = None, skip_compute: bool = False):
    """model_skip_mod_globalgate.py:255-322 (SkipGateESANet.forward).

    Returns a dict: ``out`` (logits, or the 4-scale tuple when training),
    ``weight`` [B,5], ``loss`` (FLOP regulariser), ``gate_logits`` [B,5] or
    None, ``fuse`` (list of the 4 fused stage outputs) and ``stem``
    (pooled rgb, pooled depth).

    ``weight`` overrides the gate (forced branches).  ``skip_compute=True``
    evaluates what the CUDA engine evaluates under hard gates: depth stage s is
    only run for samples whose one-hot branch index is >= s (everything else is
    identical) -- used by the skip-equivalence tests.
    """
    c = _Ctx(sd, training, cfg.activation)
    se = cfg.fuse_depth_in_rgb_encoder == "SE-add"
    r = encoder_first_conv(c, "encoder_rgb", rgb)
    d = encoder_first_conv(c, "encoder_depth", depth)
    fuse = se_fusion_add(c, "se_layer0", r, d) if se else r + d
    r = F.max_pool2d(fuse, 3, 2, 1)
    d = F.max_pool2d(d, 3, 2, 1)
    stem = (r, d)
    bs = r.shape[0]
    logits = None
    if weight is not None:
        pass
    elif baseline:
        weight = torch.zeros(bs, 5)
        weight[:, 4] = 1
    elif ini_stage:
        weight = torch.zeros(bs, 5)
        idx = torch.randint(0, 5, (bs,))          # global CPU generator, :269
        weight[range(bs), idx] = 1
    else:
        logits = global_gate_logits(c, r, d)
        weight = diff_softmax(logits, temp, hard_gate, 1)

    fused = []
    skips = []
    if not skip_compute:
        for s in (1, 2, 3, 4):
            r = encoder_layer(c, cfg, "encoder_rgb", s, r if s == 1 else fuse)
            d = encoder_layer(c, cfg, "encoder_depth", s, d)
            b0 = r
            b1 = se_fusion_add(c, f"se_layer{s}", r, d) if se else r + d
            if s < 4:
                w = weight[:, :s].sum(1).view(-1, 1, 1, 1)
                fuse = w * b0 + (1 - w) * b1
            else:
                w = weight[:, 4].view(-1, 1, 1, 1)
                fuse = (1 - w) * b0 + w * b1
            fused.append(fuse)
    else:
        assert not training, "skipping is an inference feature (batch-norm couples samples in training)"
        branch = weight.argmax(1)
        for s in (1, 2, 3, 4):
            r = encoder_layer(c, cfg, "encoder_rgb", s, r if s == 1 else fuse)
            keep = (branch >= s).nonzero().flatten()
            fuse = r.clone()
            if keep.numel():
                d_k = encoder_layer(c, cfg, "encoder_depth", s, d[keep])
                d = d.new_zeros((bs,) + d_k.shape[1:])
                d[keep] = d_k
                if se:
                    fuse[keep] = se_fusion_add(c, f"se_layer{s}", r[keep], d_k)
                else:
                    fuse[keep] = r[keep] + d_k
            fused.append(fuse)
    for s in (1, 2, 3):
        k = f"skip_layer{s}.0"
        skips.append(conv_bn_act(c, k, fused[s - 1], 1) if k + ".conv.weight" in sd else fused[s - 1])
    ctx = context_ppm(c, fused[3], cfg.upsampling) if "ppm" in cfg.context_module else fused[3]
    out = decoder(c, cfg, [ctx, skips[2], skips[1], skips[0]])
    table = DEPTH_ENC_FLOP_R34 if cfg.encoder == "resnet34" else DEPTH_ENC_FLOP_OTHER
    loss = (weight.mean(0) * torch.tensor(table)).mean()
    return {"out": out, "weight": weight, "loss": loss, "gate_logits": logits, "fuse": fused, "stem": stem}


# --------------------------------------------------------------------------
# conv MAC counter: re-derives the reference's FLOP tables (known answers)
# --------------------------------------------------------------------------

def conv_macs(cfg: FusionConfig) -> Dict[str, float]:
    """Analytic multiply-accumulate counts per image (conv layers only) split
    the way the reference's tables are (model_skip_mod_globalgate.py:419-424):
    depth stem, depth stages 1-4, and everything else."""
    H, W = cfg.height, cfg.width
    specs = {k: s for k, s, kind in _param_specs(cfg) if kind in ("conv", "upw")}

    def macs(key, h, w):
        cout, cin, kh, kw = specs[key + ".weight"]
        return cout * cin * kh * kw * h * w

    res = {"depth_stem": macs("encoder_depth.conv1", H // 2, W // 2),
           "rgb_stem": macs("encoder_rgb.conv1", H // 2, W // 2)}
    for enc in ("encoder_rgb", "encoder_depth"):
        for s in range(4):
            h, w = H // (4 << s), W // (4 << s)
            tot = 0
            for k in specs:
                k = k[:-len(".weight")]
                if not k.startswith(f"{enc}.layer{s + 1}."):
                    continue
                # the strided 3x1 of NBt1D reduces H only; its output is (h, 2w)
                if k.endswith(".0.conv3x1_1") and s > 0:
                    tot += macs(k, h, 2 * w)
                elif cfg.block == "Bottleneck" and k.endswith(".0.conv1") and s > 0:
                    tot += macs(k, 2 * h, 2 * w)
                else:
                    tot += macs(k, h, w)
            res[f"{enc}.stage{s + 1}"] = tot
    return res
