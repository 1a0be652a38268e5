"""Inference engine of the gated RGB-D encoder/decoder on the CUDA kernels.

Replaces the body of ``SkipGateESANet.forward`` in eval mode
(model_skip_mod_globalgate.py:255-322 of the reference).  Weight preparation
(BN folding, bf16 K-major repacking) happens once per ``state_dict``; the
forward is a fixed sequence of C-ABI launches whose data-dependent parts (which
samples run which depth stage) are resolved on the device from the gate's
output, so the whole forward is CUDA-graph capturable and never syncs the host.

Real skipping: depth samples are kept in *slot order* (sorted by how many depth
stages their branch needs), so the set of samples a depth stage must process is
always a prefix ``[0, count[s])`` -- the conv kernels size their tile lists from
``count[s]`` on the device, and the RGB stage's last conv only reads the depth
features of samples whose gate is non-zero.  The RGB and depth encoders run on
two streams and meet at the four fusion points.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Callable, Dict, List, Optional, Sequence

import torch

from .. import _lib, ops

Tensor = torch.Tensor

_STAGE_BLOCKS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}
_PLANES = (64, 128, 256, 512)


def pack_gate(sd: Dict[str, Tensor], prefix: str = "gate_layer.") -> Dict[str, Tensor]:
    """GlobalGate parameters (model_skip_mod_globalgate.py:379-386) in the layout
    dynmm_global_gate_logits expects; conv bias + eval BN folded to scale/shift."""
    g = lambda k: sd[prefix + k].detach().float()
    cuda = g("conv.0.weight").is_cuda
    fold = ops.fold_bn_cuda if cuda else ops.fold_bn
    s1, b1 = fold(g("conv.1.weight"), g("conv.1.bias"), g("conv.1.running_mean"), g("conv.1.running_var"),
                  1e-5, g("conv.0.bias"))
    s2, b2 = fold(g("conv.4.weight"), g("conv.4.bias"), g("conv.4.running_mean"), g("conv.4.running_var"),
                  1e-5, g("conv.3.bias"))

    def taps_last_channel(w):          # [o][c][kh][kw] -> [o][kh][kw][c]
        o, c, kh, kw = w.shape
        if cuda:
            return ops.permute3d(w.reshape(o, c, kh * kw), (0, 2, 1)).view(o, kh, kw, c)
        return w.permute(0, 2, 3, 1).contiguous()
    return {
        "w1": taps_last_channel(g("conv.0.weight")),                 # [8][5][5][128]
        "s1": s1, "b1": b1,
        "w2": taps_last_channel(g("conv.3.weight")),                 # [8][5][5][8]
        "s2": s2, "b2": b2,
        "wfc": g("fc.weight").reshape(g("fc.weight").shape[0], -1).contiguous(),
    }


@dataclasses.dataclass
class ConvLayer:
    """One convolution with everything that follows it folded in."""
    weight: Tensor                 # packed bf16 [taps, c_out_pad, c_in]
    scale: Optional[Tensor]
    shift: Optional[Tensor]
    c_in: int
    c_out: int
    kh: int
    kw: int
    stride: tuple
    pad: tuple
    relu: int                      # activation code of dynmm_conv_params.relu (0 none, 1 ReLU, 2 swish, 3 h-swish)
    split: bool = False            # fp32-grade mode: [hi | lo] activations, weight packed [W_hi | W_lo | W_hi]

    def __call__(self, x: Tensor, **kw) -> Tensor:
        return ops.conv(x, self.weight, c_out=self.c_out, kh=self.kh, kw=self.kw, stride=self.stride, pad=self.pad,
                        scale=self.scale, shift=self.shift, relu=self.relu, c_in=self.c_in, split=self.split, **kw)


@dataclasses.dataclass
class Block:
    convs: List[ConvLayer]          # main path; the last one takes the residual
    downsample: Optional[ConvLayer]


class _Packer:
    def __init__(self, sd: Dict[str, Tensor], device, split: bool = False, act: int = 1):
        self.sd, self.dev, self.split, self.act = sd, device, split, act

    def t(self, key):
        return self.sd[key].detach().float().to(self.dev)

    def taps_major(self, key):
        """depthwise 3x3 stencil [C][1][3][3] -> tap-major [9][C] (dynmm_upsample2x_dw3x3)."""
        w = self.t(key)
        return ops.permute3d(w.reshape(w.shape[0], 1, 9), (2, 1, 0)).view(9, w.shape[0])

    def conv(self, key, *, stride=(1, 1), pad=(0, 0), bn: Optional[str] = None, bn_eps=1e-5, relu=False,
             pad_out: int = 1) -> ConvLayer:
        """BN scale folded into the fp32 weights before bf16 packing, the epilogue only adds `shift`: one launch of
        dynmm_fold_pack_conv per convolution (no library element-wise kernels in the engine build)."""
        w = self.t(key + ".weight")
        bias = self.t(key + ".bias") if key + ".bias" in self.sd else None
        bn_t = [self.t(bn + s) for s in (".weight", ".bias", ".running_mean", ".running_var")] if bn is not None else None
        packed, shift = ops.fold_pack_conv(w, bias, bn_t, bn_eps, split=self.split)
        c_out, c_in, kh, kw = w.shape
        if c_out % pad_out:
            # extra output channels: the packed weight already has zero rows up to a multiple of 16; zero shifts
            c_pad = (c_out + pad_out - 1) // pad_out * pad_out
            if shift is not None:
                shift = torch.cat([shift, torch.zeros(c_pad - c_out, device=shift.device)]).contiguous()
            c_out = c_pad
        # `relu=True` means "the model's activation": its code for dynmm_conv_params.relu
        return ConvLayer(packed, None, shift, c_in, c_out, kh, kw, tuple(stride), tuple(pad), self.act if relu else 0,
                         self.split)

    def nbt1d(self, key, stride=1) -> Block:
        """resnet.py:124-147: 3x1 -> ReLU -> 1x3 -> BN(1e-3) -> ReLU -> 3x1 -> ReLU -> 1x3 -> BN -> +id -> ReLU."""
        convs = [
            self.conv(key + ".conv3x1_1", stride=(stride, 1), pad=(1, 0), relu=True),
            self.conv(key + ".conv1x3_1", stride=(1, stride), pad=(0, 1), bn=key + ".bn1", bn_eps=1e-3, relu=True),
            self.conv(key + ".conv3x1_2", pad=(1, 0), relu=True),
            self.conv(key + ".conv1x3_2", pad=(0, 1), bn=key + ".bn2", bn_eps=1e-3, relu=True),
        ]
        return Block(convs, self._downsample(key, stride))

    def basic(self, key, stride=1) -> Block:
        """resnet.py:66-84."""
        convs = [
            self.conv(key + ".conv1", stride=(stride, stride), pad=(1, 1), bn=key + ".bn1", relu=True),
            self.conv(key + ".conv2", pad=(1, 1), bn=key + ".bn2", relu=True),
        ]
        return Block(convs, self._downsample(key, stride))

    def bottleneck(self, key, stride=1) -> Block:
        """resnet.py:173-192."""
        convs = [
            self.conv(key + ".conv1", bn=key + ".bn1", relu=True),
            self.conv(key + ".conv2", stride=(stride, stride), pad=(1, 1), bn=key + ".bn2", relu=True),
            self.conv(key + ".conv3", bn=key + ".bn3", relu=True),
        ]
        return Block(convs, self._downsample(key, stride))

    def _downsample(self, key, stride):
        if key + ".downsample.0.weight" not in self.sd:
            return None
        return self.conv(key + ".downsample.0", stride=(stride, stride), bn=key + ".downsample.1")

    def conv_bn_act(self, key, k) -> ConvLayer:
        return self.conv(key + ".conv", pad=(k // 2, k // 2), bn=key + ".bn", relu=True)


@dataclasses.dataclass
class EngineConfig:
    encoder: str = "resnet34"
    encoder_depth: Optional[str] = None     # None: the RGB encoder's architecture (model_skip_mod_globalgate.py:81-93)
    encoder_block: str = "NonBottleneck1D"
    fuse: str = "add"
    nr_decoder_blocks: Sequence[int] = (3, 3, 3)
    num_classes: int = 40
    upsampling: str = "learned-3x3-zeropad"
    context_module: str = "ppm"
    activation: str = "relu"
    gate: str = "global"            # "global": GlobalGate (SkipGateESANet); "local": one two-way gate per fusion site
    # "bf16": bf16 activations and weights, fp32 accumulation (stated tolerance 2e-2 on the logits);
    # "f32x3": fp32-grade arithmetic on the same tensor cores -- activations and weights as bf16 hi + lo halves, three
    # products per MAC (DYNMM_CONV_SPLIT), element-wise kernels in fp32: logits within 1e-3 of the fp32 reference
    precision: str = "bf16"
    encoder_decoder_fusion: str = "add"     # "None": the decoder modules do not add the encoder skip tensors (model.py:353-355)


class FusionEngine:
    """Packed weights + launch sequence for one ``state_dict``."""

    def __init__(self, sd: Dict[str, Tensor], cfg: EngineConfig, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DynmmError("FusionEngine needs a CUDA device")
        with torch.cuda.device(device):          # packing kernels, streams and kernel attributes are per device
            self._build(sd, cfg, device)

    def _build(self, sd: Dict[str, Tensor], cfg: EngineConfig, device):
        _lib.require_device()
        if cfg.activation.lower() not in ("relu", "swish", "silu", "hswish"):
            raise NotImplementedError("unknown activation " + cfg.activation)
        # ReLU is the tuned path.  swish / h-swish (model_utils.py:100-115) run through the same convolution kernel
        # (activation code in the epilogue); the fused 64-channel pair kernel, the chain kernel and the tensor-core
        # stem hard-wire ReLU, so those models use per-layer launches and a library (cuDNN) stem.
        self.act = ops.ACTIVATIONS[cfg.activation.lower()]
        if self.act != 1 and (cfg.fuse != "add" or cfg.gate != "global"):
            raise NotImplementedError("swish / h-swish are implemented for the global gate with fuse='add' (the SE "
                                      "kernels and the local gates' squeeze-excite hard-wire ReLU)")
        if cfg.fuse not in ("add", "SE-add"):
            raise NotImplementedError("fuse_depth_in_rgb_encoder must be 'add' or 'SE-add', got " + str(cfg.fuse))
        if cfg.upsampling not in ("learned-3x3-zeropad", "learned-3x3", "bilinear", "nearest"):
            raise NotImplementedError("unknown upsampling mode " + str(cfg.upsampling))
        if "ppm" not in cfg.context_module or "appm" in cfg.context_module:
            raise NotImplementedError("the CUDA engine implements context_module='ppm' (bins 1,5) and 'ppm-1-2-4-8'")
        if cfg.precision not in ("bf16", "f32x3"):
            raise ValueError("EngineConfig.precision must be 'bf16' or 'f32x3'")
        self.split = cfg.precision == "f32x3"
        if self.split and cfg.activation.lower() != "relu":
            raise NotImplementedError("precision='f32x3' is implemented for ReLU models")
        if cfg.encoder_decoder_fusion not in ("add", "None"):
            raise NotImplementedError("encoder_decoder_fusion must be 'add' or 'None'")
        self.dec_fusion = cfg.encoder_decoder_fusion == "add"
        self.cfg, self.dev = cfg, device
        p = _Packer(sd, device, self.split, self.act)
        if cfg.gate == "local":
            if cfg.fuse != "add":
                raise NotImplementedError("the local-gate engine blends by addition (model_skip_mod.py:241-311)")
            # SqueezeAndExciteReweigh of site i (rgb_depth_fusion.py:29-65): the SE 1x1 convs as fp32 matrices
            self.gate = None
            self.local_gates = []
            for i in range(4):
                k = f"gate_layer{i}.se.fc"
                w1, w2 = p.t(k + ".0.weight"), p.t(k + ".2.weight")
                self.local_gates.append((w1.reshape(w1.shape[0], -1).contiguous(), p.t(k + ".0.bias").contiguous(),
                                         w2.reshape(w2.shape[0], -1).contiguous(), p.t(k + ".2.bias").contiguous()))
        else:
            self.gate = {k: v.to(device) for k, v in pack_gate(sd).items()}
        # stem: [7][7][cin][64] fp32 + folded BN
        self.stem = {}
        for enc in ("encoder_rgb", "encoder_depth"):
            w0 = p.t(enc + ".conv1.weight")                               # [64][cin][7][7] -> [7][7][cin][64]
            w = ops.permute3d(w0.reshape(w0.shape[0], w0.shape[1], 49), (2, 1, 0)).view(7, 7, w0.shape[1], w0.shape[0])
            s, b = ops.fold_bn_cuda(p.t(enc + ".bn1.weight"), p.t(enc + ".bn1.bias"), p.t(enc + ".bn1.running_mean"),
                                    p.t(enc + ".bn1.running_var"), 1e-5)
            self.stem[enc] = (w, s, b)
        # plain `add` fusion: TMA-gathered stem (dynmm_stem_s2d_fwd); DYNMM_STEM=tc|fp32 selects the older kernels
        self.stem_packed = None
        if cfg.fuse == "add" and os.environ.get("DYNMM_STEM", "s2d") == "s2d" and self.act == 1:
            self.stem_packed = ops.stem_s2d_pack_weights(self.stem["encoder_rgb"][0], self.stem["encoder_depth"][0])
            # host copy of the BN vectors: they travel as a kernel parameter (constant bank) on every stem launch
            self.stem_bn_host = ops.stem_s2d_bn_host(*self.stem["encoder_rgb"][1:], *self.stem["encoder_depth"][1:])
        # the two encoders may differ (e.g. ResNet-34 for RGB, ResNet-18 for depth); their stage outputs must have the
        # same channel counts, because every fusion site adds them
        enc_arch = {"encoder_rgb": cfg.encoder, "encoder_depth": cfg.encoder_depth or cfg.encoder}
        self.stages = {}
        for enc, arch in enc_arch.items():
            block_kind = "bottleneck" if arch == "resnet50" else \
                {"NonBottleneck1D": "nbt1d", "BasicBlock": "basic"}[cfg.encoder_block]
            make = getattr(p, block_kind)
            stages = []
            for s, nblk in enumerate(_STAGE_BLOCKS[arch]):
                stages.append([make(f"{enc}.layer{s + 1}.{b}", 2 if (b == 0 and s > 0) else 1) for b in range(nblk)])
            self.stages[enc] = stages
        self.same_encoders = enc_arch["encoder_rgb"] == enc_arch["encoder_depth"]
        if [st[-1].convs[-1].c_out for st in self.stages["encoder_rgb"]] != \
                [st[-1].convs[-1].c_out for st in self.stages["encoder_depth"]]:
            raise NotImplementedError("the encoders' stage outputs must have equal channel counts (they are added)")
        self.stage_channels = [st[-1].convs[-1].c_out for st in self.stages["encoder_rgb"]]
        self.skips = [p.conv_bn_act(f"skip_layer{i}.0", 1) if f"skip_layer{i}.0.conv.weight" in sd else None
                      for i in (1, 2, 3)]
        # pyramid pooling bins (context_modules.py:28-38): 'ppm' = (1, 5), 'ppm-1-2-4-8' = (1, 2, 4, 8)
        self.ppm_bins = (1, 2, 4, 8) if cfg.context_module == "ppm-1-2-4-8" else (1, 5)
        self.ppm = [p.conv_bn_act(f"context_module.features.{i}.1", 1) for i in range(len(self.ppm_bins))]
        self.ppm_c = sum(cv.c_out for cv in self.ppm)          # channels the branches add to the concat buffer
        self.ppm_final = p.conv_bn_act("context_module.final_conv", 1)
        self.dec = []
        for i in range(3):
            dk = f"decoder.decoder_module_{i + 1}"
            self.dec.append({
                "conv3x3": p.conv_bn_act(dk + ".conv3x3", 3),
                "blocks": [p.nbt1d(f"{dk}.decoder_blocks.{b}") for b in range(cfg.nr_decoder_blocks[i])],
                "up_w": None, "up_b": None,
            })
            self.dec[-1]["up_w"], self.dec[-1]["up_b"] = self._up_params(p, dk + ".upsample", self.dec[-1]["conv3x3"].c_out)
        # NHWC kernels move 16-byte channel groups: a class count that is not a multiple of 8 (SUN RGB-D: 37) is carried
        # as the next multiple (zero weight rows / stencils / biases); the final kernel emits the real classes only
        self.n_classes = cfg.num_classes
        self.conv_out = p.conv("decoder.conv_out", pad=(1, 1), pad_out=8)
        self.up = [self._up_params(p, f"decoder.{u}", self.conv_out.c_out) for u in ("upsample1", "upsample2")]
        # model.py:360-410: 'learned-3x3' pads the up-sampled map by replication, 'bilinear' is that form with the
        # fixed [1 2 1]^T [1 2 1] / 16 stencil, 'nearest' the identity stencil; the context module interpolates its
        # pooled branches bilinearly only for 'bilinear' (model_skip_mod_globalgate.py:145-160: learned -> nearest)
        self.up_replicate = cfg.upsampling in ("learned-3x3", "bilinear")
        self.ctx_bilinear = cfg.upsampling == "bilinear"
        # SE-add fusion: the 1x1 convs of SqueezeAndExcitation as fp32 matrices (model_utils.py:40-45)
        self.se = None
        if cfg.fuse == "SE-add":
            self.se = []
            for i in range(5):
                layer = {}
                for stream in ("se_rgb", "se_depth"):
                    k = f"se_layer{i}.{stream}.fc"
                    w1 = p.t(k + ".0.weight")
                    w2 = p.t(k + ".2.weight")
                    layer[stream] = (w1.reshape(w1.shape[0], -1).contiguous(), p.t(k + ".0.bias").contiguous(),
                                     w2.reshape(w2.shape[0], -1).contiguous(), p.t(k + ".2.bias").contiguous())
                self.se.append(layer)
        self.side = torch.cuda.Stream(device=device)
        self.side2 = torch.cuda.Stream(device=device)      # skip-connection 1x1 convs, off the encoders' critical path
        self.launches = 0          # kernels launched by the last forward (for bench accounting)
        # DYNMM_PROGRAM=1 ('add' fusion only): the encoders, the skip convs and each decoder module run as persistent
        # convolution PROGRAMS (one cooperative launch per chain of dependent convs, dynmm_conv_program_*) instead of
        # one launch per convolution on two streams.  Same arithmetic, bit-identical results; measured slower at
        # batch 8 in round 1 (profiles/r1_program_*), so the per-launch path stays the default.
        self.use_programs = (cfg.fuse == "add" and os.environ.get("DYNMM_PROGRAM", "0") == "1" and not self.split and
                             self.dec_fusion and self.same_encoders and self.act == 1)
        self.programs: list = []   # ConvPrograms of the last forward (a captured graph must keep them alive)
        # 64-channel NonBottleneck1D blocks: each 3x1 -> 1x3 pair as ONE fused kernel (dynmm_conv_pair_fwd, bit-identical
        # to the two launches); DYNMM_PAIR=0 keeps one launch per convolution
        self.use_pairs = os.environ.get("DYNMM_PAIR", "1") != "0" and not self.split and self.act == 1
        # DYNMM_TILE_FLAGS=1: convolutions publish per-tile completion flags and their consumers wait on those instead
        # of on the previous kernel as a whole (layer k+1 starts on the SMs layer k's early finishers free)
        self.flag_pool = ops.TileFlagPool(device) if (os.environ.get("DYNMM_TILE_FLAGS", "0") == "1" and
                                                      not self.split) else None
        # DYNMM_MERGE (default on; 'add' fusion): from stage 2 on, the same layer of the RGB and of the depth encoder is ONE
        # launch (dynmm_conv_igemm_fwd2) on one stream; only the last convolution of a stage runs per encoder (the RGB
        # one adds g_s * depth_s, which the depth one has to finish first).  Stage 1 (64 channels: fused-pair kernels
        # with resident weights) keeps the two-stream form.
        self.use_merge = (cfg.fuse == "add" and os.environ.get("DYNMM_MERGE", "1") == "1" and not self.use_programs and
                          self.same_encoders)
        # DYNMM_CHAIN (default on, merged stages only): blocks 2.. of an encoder stage (stride 1, 128 or 256 channels) run
        # as ONE chain kernel per stage for both encoders (dynmm_conv_chain_fwd: activations stay in shared memory from
        # layer to layer); the stage's first block (stride 2, down-sampling) and the RGB encoder's last convolution (gated
        # add) stay per-layer launches.  DYNMM_CHAIN_STAGES: comma-separated stage indices (0-based), default "2".
        self.chain_imgs = {}
        self._num_sms = torch.cuda.get_device_properties(device).multi_processor_count
        self._may_skip = True
        if self.use_merge and os.environ.get("DYNMM_CHAIN", "1") == "1" and not self.split and self.act == 1:
            for s in {int(v) for v in os.environ.get("DYNMM_CHAIN_STAGES", "2").split(",") if v.strip()}:
                if 1 <= s <= 3:
                    imgs = self._chain_images(s)
                    if imgs is not None:
                        self.chain_imgs[s] = imgs

    def _library_stem(self, rgb: Tensor, depth: Tensor):
        """resnet.py:352-358 + model_skip_mod_globalgate.py:256-261 for the activations the tensor-core stem does not
        hard-wire (swish / h-swish): cuDNN convolution, folded BatchNorm, activation, add, max-pool, NHWC copies.
        -> (rgb fp32 NHWC, depth fp32 NHWC, rgb bf16 NHWC, depth bf16 NHWC), pooled to 1/4 resolution."""
        import torch.nn.functional as F

        def act(x):
            return x * torch.sigmoid(x) if self.act == 2 else x * F.relu6(x + 3.0) / 6.0

        outs = []
        for x, enc in ((rgb, "encoder_rgb"), (depth, "encoder_depth")):
            w, s, b = self.stem[enc]                                   # [7][7][cin][64], folded BN scale / shift
            y = F.conv2d(x, w.permute(3, 2, 0, 1).contiguous(), stride=2, padding=3)
            outs.append(act(y * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)))
        r = F.max_pool2d(outs[0] + outs[1], 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        d = F.max_pool2d(outs[1], 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        return r, d, r.to(torch.bfloat16), d.to(torch.bfloat16)

    def _up_params(self, p: "_Packer", key: str, c: int):
        """(tap-major [9][c] stencil, bias or None) of one Upsample module for dynmm_upsample2x_dw3x3."""
        mode = self.cfg.upsampling
        if "learned-3x3" in mode:
            w, b = p.taps_major(key + ".conv.weight"), p.t(key + ".conv.bias").contiguous()
            if w.shape[1] < c:                       # padded class channels: zero stencil, zero bias
                w = torch.cat([w, torch.zeros(9, c - w.shape[1], device=w.device)], dim=1).contiguous()
                b = torch.cat([b, torch.zeros(c - b.shape[0], device=b.device)]).contiguous()
            return w, b
        if mode == "bilinear":
            return ops.bilinear_stencil(c, self.dev), None
        return ops.nearest_stencil(c, self.dev), None

    # ------------------------------------------------------------------ blocks
    @staticmethod
    def _pairable(blk: Block) -> bool:
        """NonBottleneck1D block whose two 3x1 -> 1x3 pairs fit dynmm_conv_pair_fwd (64 channels, stride 1)."""
        if blk.downsample is not None or len(blk.convs) != 4:
            return False
        shapes = [(3, 1), (1, 3), (3, 1), (1, 3)]
        return all(c.c_in == 64 and c.c_out == 64 and (c.kh, c.kw) == k and c.stride == (1, 1) and c.scale is None
                   and c.relu for c, k in zip(blk.convs, shapes))

    def _block(self, x: Tensor, blk: Block, keep: list, *, count=None, in_map=None, before_last: Callable = None,
               last_kw: Optional[dict] = None, count_settled: bool = False) -> Tensor:
        """``count_settled``: the gate plan that wrote ``count`` completed before the previous kernel of this stream
        started (every depth block but the first) -- the conv kernels then read it ahead of their launch wait."""
        n_out = x.shape[0]
        cs = dict(count_settled=True) if (count_settled and count is not None) else {}
        if self.use_pairs and self._pairable(blk):
            c0, c1, c2, c3 = blk.convs
            y = ops.conv_pair(x, c0.weight, c0.shift, c1.weight, c1.shift, relu2=True, count=count, in_map=in_map,
                              n_out=n_out)
            keep.append(y)
            self.launches += 1
            if before_last is not None:
                before_last()
            if last_kw:                # gated add / redirected output: the second pair stays two launches
                y = c2(y, count=count, n_out=n_out, **cs)
                out = c3(y, residual=x, res_map=in_map, count=count, n_out=n_out, **cs, **last_kw)
                keep += [y, out]
                self.launches += 2
            else:
                out = ops.conv_pair(y, c2.weight, c2.shift, c3.weight, c3.shift, residual=x, res_map=in_map, relu2=True,
                                    count=count, n_out=n_out)
                keep.append(out)
                self.launches += 1
            return out
        y = x
        for i, cv in enumerate(blk.convs[:-1]):
            y = cv(y, count=count, in_map=in_map if i == 0 else None, n_out=n_out, **cs)
            keep.append(y)
            self.launches += 1
        if blk.downsample is not None:
            idn = blk.downsample(x, count=count, in_map=in_map, n_out=n_out, **cs)
            keep.append(idn)
            self.launches += 1
            res_map = None
        else:
            idn, res_map = x, in_map
        if before_last is not None:
            before_last()
        first = len(blk.convs) == 1
        # residual_settled: a residual without flags is the block input, which the block's first convolution (an
        # ordinary stream-ordered launch in that case) has already waited for
        out = blk.convs[-1](y, residual=idn, res_map=res_map, count=count, in_map=in_map if first else None,
                            n_out=n_out, residual_settled=(blk.downsample is None and not first), **cs,
                            **(last_kw or {}))
        keep.append(out)
        self.launches += 1
        return out

    def _chain_images(self, s: int):
        """ChainImages (RGB without its last convolution, depth complete) for blocks 1.. of stage s, or None when the
        blocks are not plain 128- / 256-channel NonBottleneck1D blocks."""
        out = []
        for enc, drop in (("encoder_rgb", True), ("encoder_depth", False)):
            blocks = self.stages[enc][s][1:]
            if not blocks:
                return None
            c = blocks[0].convs[0].c_in
            shapes = [(3, 1), (1, 3), (3, 1), (1, 3)]
            for blk in blocks:
                if blk.downsample is not None or len(blk.convs) != 4 or c not in (128, 256):
                    return None
                if not all(cv.c_in == c and cv.c_out == c and (cv.kh, cv.kw) == k and cv.stride == (1, 1) and
                           cv.scale is None and cv.pad == (k[0] // 2, k[1] // 2) for cv, k in zip(blk.convs, shapes)):
                    return None
            layers = ops.nbt1d_chain_layers([[(cv.weight, cv.shift, cv.relu) for cv in blk.convs] for blk in blocks],
                                            drop_last=drop)
            out.append(ops.ChainImage(layers, c, self.dev))
        return tuple(out)

    def _chain_flags(self, n: int) -> Tensor:
        """n zeroed int32 flags for one chain launch, carved from the pool this forward zeroed at its start."""
        if self._flag_off + n > self._flag_pool.numel():
            return torch.zeros(n, dtype=torch.int32, device=self.dev)
        f = self._flag_pool[self._flag_off:self._flag_off + n]
        self._flag_off += (n + 3) // 4 * 4
        return f

    def _merged_stage(self, s: int, r: Tensor, d: Tensor, plan, keep: list, cat: Optional[Tensor]):
        """Stage s (>= 1) of both encoders in lock step: layer i of the depth encoder (slot order, prefix-counted)
        and layer i of the RGB encoder share a launch; the stage's last convolution runs per encoder -- depth
        first, then the RGB one with the gated add.  -> (fused RGB stage output, depth stage output)"""
        cnt = plan.count[s:s + 1]
        dk = dict(count=cnt, count_settled=True)
        blocks_r, blocks_d = self.stages["encoder_rgb"][s], self.stages["encoder_depth"][s]
        n_d = d.shape[0]
        chain = self.chain_imgs.get(s)
        if chain is not None:
            # geometry of blocks 1..: the stage's first block halves the map
            ho, wo = (r.shape[1] + 1) // 2, (r.shape[2] + 1) // 2
            cplan = ops.chain_plan(ho, wo, chain[0].c, r.shape[0] + n_d)
            # The strips of a sample wait for each other, so a chain launch wants all its CTAs resident at once.  The
            # RGB job always has n strips-sets; the depth job has `count` of them, known on the device only.  When even
            # the worst case fits (every depth sample active) the chain always wins; otherwise it wins as long as the
            # gate actually skips (hard decisions: measured 208 us vs 266 us per stage at 4 of 8 samples, but 398 us vs
            # ~300 us with all 8 active) -- soft gates and the baseline mode never skip, so they keep per-layer launches.
            if cplan is None or (cplan[0] > self._num_sms and not self._may_skip):
                chain = None
        for bi, (br, bd) in enumerate(zip(blocks_r, blocks_d)):
            if chain is not None and bi == 1:
                (out_r, last_r), (d, _) = ops.conv_chain(
                    [dict(x=r, image=chain[0]), dict(x=d, image=chain[1], **dk)],
                    flags=self._chain_flags(cplan[0] + r.shape[0] + n_d))
                kw = dict(gated=d, gate=plan.g[s], gated_slot=plan.slot)
                if cat is not None:
                    kw.update(out=cat, out_c_off=0)
                keep += [r, out_r, last_r, d]
                r = blocks_r[-1].convs[-1](last_r, residual=out_r if out_r is not None else r, **kw)
                keep.append(r)
                self.launches += 2
                break
            last_block = bi == len(blocks_r) - 1
            yr, yd = r, d
            for i in range(len(br.convs) - 1):
                with ops.ConvMerge():
                    yr = br.convs[i](yr)
                    yd = bd.convs[i](yd, n_out=n_d, **dk)
                keep += [yr, yd]
                self.launches += 1
            idn_r, idn_d = r, d
            if br.downsample is not None:
                with ops.ConvMerge():
                    idn_r = br.downsample(r)
                    idn_d = bd.downsample(d, n_out=n_d, **dk)
                keep += [idn_r, idn_d]
                self.launches += 1
            if not last_block:
                with ops.ConvMerge():
                    r = br.convs[-1](yr, residual=idn_r)
                    d = bd.convs[-1](yd, residual=idn_d, n_out=n_d, **dk)
                self.launches += 1
            else:
                d = bd.convs[-1](yd, residual=idn_d, n_out=n_d, **dk)
                kw = dict(gated=d, gate=plan.g[s], gated_slot=plan.slot)
                if cat is not None:
                    kw.update(out=cat, out_c_off=0)
                r = br.convs[-1](yr, residual=idn_r, **kw)
                self.launches += 2
            keep += [r, d]
        return r, d

    def _se_fuse(self, s: int, rgb: Tensor, depth: Tensor, plan, keep: list, out: Optional[Tensor]) -> Tensor:
        """fuse = w*rgb + (1-w)*se_layer{s+1}(rgb, depth)  (model_skip_mod_globalgate.py:280-283)."""
        layer = self.se[s + 1]
        n, hh, ww, c = rgb.shape
        inv_area = 1.0 / (hh * ww)
        pr = ops.gap_partial(rgb, split=self.split)
        pd = ops.gap_partial(depth, count=plan.count[s:s + 1], split=self.split)
        sig_r = ops.se_mlp(pr, inv_area, *layer["se_rgb"])
        sig_d = ops.se_mlp(pd, inv_area, *layer["se_depth"], count=plan.count[s:s + 1])
        fused = ops.se_gated_fuse(rgb, depth, sig_r, sig_d, plan.g[s], plan.slot, out=out, split=self.split)
        keep += [pr, pd, sig_r, sig_d, fused]
        self.launches += 5
        return fused


    # ------------------------------------------------------------------ convolution programs
    def _block_steps(self, x: Tensor, blk: Block, keep: list, result: list, *, count=None, in_map=None,
                     last_kw: Optional[dict] = None):
        """Generator form of :meth:`_block`: records one phase worth of convolutions per step (the first
        conv together with the block's down-sampling conv), the block output is appended to ``result``."""
        n_out = x.shape[0]
        assert len(blk.convs) >= 2 or blk.downsample is None
        y, idn, res_map = x, x, in_map
        for i, cv in enumerate(blk.convs):
            kw = dict(count=count, in_map=in_map if i == 0 else None, n_out=n_out)
            if i == 0 and blk.downsample is not None:
                idn, res_map = blk.downsample(x, count=count, in_map=in_map, n_out=n_out), None
                keep.append(idn)
            if i == len(blk.convs) - 1:
                # evaluated only now: the depth features a gated add reads were recorded one phase earlier
                kw.update(residual=idn, res_map=res_map, **(last_kw() if callable(last_kw) else (last_kw or {})))
            y = cv(y, **kw)
            keep.append(y)
            yield
        result.append(y)

    def _encoder_chain(self, x: Tensor, enc: str, keep: list, outs: list, plan, *, depth_outs=None, cat=None):
        """All blocks of one encoder as a generator of phases.  Depth encoder (``depth_outs is None``): slot
        order, prefix-counted.  RGB encoder: the last conv of stage s adds g_s * depth_s."""
        is_depth = depth_outs is None
        for s in range(4):
            blocks = self.stages[enc][s]
            for bi, blk in enumerate(blocks):
                kw, last_kw = {}, None
                if is_depth:
                    kw = dict(count=plan.count[s:s + 1], in_map=plan.perm if (s == 0 and bi == 0) else None)
                elif bi == len(blocks) - 1:
                    def last_kw(s=s):
                        kw_ = dict(gated=depth_outs[s], gate=plan.g[s], gated_slot=plan.slot)
                        if s == 3:
                            kw_.update(out=cat, out_c_off=0)
                        return kw_
                res: list = []
                yield from self._block_steps(x, blk, keep, res, last_kw=last_kw, **kw)
                x = res[0]
            outs.append(x)

    def _encoder_program(self, r16: Tensor, d16: Tensor, plan, cat: Tensor, keep: list):
        """Both encoders and the skip 1x1 convs as ONE program.  Phase k holds depth conv k and RGB conv k-1:
        the depth encoder runs one layer ahead, so depth_s is complete when the RGB stage's last conv reads it."""
        depth_out, fused, skips = [], [], []
        with ops.ConvProgram() as prog:
            dgen = self._encoder_chain(d16, "encoder_depth", keep, depth_out, plan)
            rgen = self._encoder_chain(r16, "encoder_rgb", keep, fused, plan, depth_outs=depth_out, cat=cat)
            d_live, r_live, first = True, True, True
            while d_live or r_live:
                if d_live:
                    d_live = next(dgen, StopIteration) is not StopIteration
                if r_live and not first:
                    r_live = next(rgen, StopIteration) is not StopIteration
                first = False
                while len(skips) < min(len(fused), 3) and prog.jobs_in_phase() < 4:
                    sk = self.skips[len(skips)]
                    skips.append(sk(fused[len(skips)]) if sk is not None else fused[len(skips)])
                prog.next_phase()
            while len(skips) < 3:
                sk = self.skips[len(skips)]
                skips.append(sk(fused[len(skips)]) if sk is not None else fused[len(skips)])
        self.programs.append(prog)
        self.launches += 2          # program image upload + the cooperative launch
        keep += skips
        return fused, skips

    def _sequence_program(self, x: Tensor, items: Sequence, keep: list) -> Tensor:
        """A chain of dependent ConvLayers / Blocks as one program (one phase per convolution)."""
        with ops.ConvProgram() as prog:
            for it in items:
                if isinstance(it, Block):
                    res: list = []
                    for _ in self._block_steps(x, it, keep, res):
                        prog.next_phase()
                    x = res[0]
                else:
                    x = it(x)
                    keep.append(x)
                    prog.next_phase()
        self.programs.append(prog)
        self.launches += 2
        return x

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, rgb: Tensor, depth: Tensor, **kw):
        """See :meth:`_forward`; runs with the engine's device current (kernel attributes, streams and launches
        are per device)."""
        with torch.cuda.device(self.dev):
            if self.flag_pool is None:
                return self._forward(rgb, depth, **kw)
            self.flag_pool.reset()                 # one memset per forward, ahead of both streams
            ops.FLAG_POOL = self.flag_pool
            try:
                return self._forward(rgb, depth, **kw)
            finally:
                ops.FLAG_POOL = None

    def _forward(self, rgb: Tensor, depth: Tensor, *, temp: float = 1.0, hard_gate: bool = False,
                 baseline: bool = False, ini_stage: bool = False, weight: Optional[Tensor] = None,
                 out: Optional[Tensor] = None, hist: Optional[Tensor] = None, labels: Optional[Tensor] = None,
                 want_logits: bool = True):
        """rgb [B,3,H,W], depth [B,1,H,W] fp32 NCHW on the GPU ->
        (logits [B,classes,H,W] fp32 NCHW, gate weight [B,5] fp32).
        ``labels`` (uint8 [B,H,W]): also emit argmax_c(logits) from the final kernel (eval.py:120);
        with ``want_logits=False`` the logits are never written and None is returned for them."""
        if not (rgb.is_cuda and depth.is_cuda):
            raise _lib.DynmmError("FusionEngine.forward needs CUDA tensors")
        rgb = rgb.float().contiguous()
        depth = depth.float().contiguous()
        b, _, h, w = rgb.shape
        if h % 32 or w % 32:
            raise _lib.DynmmError(f"input size {h}x{w} must be a multiple of 32 (five stride-2 stages)")
        self.launches = 0
        keep: list = []
        main = torch.cuda.current_stream()
        side = self.side
        if self.chain_imgs:
            self._flag_pool = torch.zeros(4096, dtype=torch.int32, device=self.dev)   # one memset per forward
            self._flag_off = 0
            keep.append(self._flag_pool)
            # can this forward skip depth samples at all?  (hard learned gate, random / forced one-hot branches)
            self._may_skip = bool(hard_gate or ini_stage or weight is not None) and not baseline
        wr, sr, br = self.stem["encoder_rgb"]
        wd, sdp, bd = self.stem["encoder_depth"]
        learned = weight is None and not baseline and not ini_stage
        if self.act != 1:
            r32, d32, r16, d16 = self._library_stem(rgb, depth)
            if self.split:
                r16, d16 = ops.split_from_f32(r32), ops.split_from_f32(d32)
            self.launches += 10
        elif self.split:
            # fp32-grade mode: the stem (three split products already) hands its fp32 maps on as [hi | lo] halves
            if self.stem_packed is not None:       # the stem writes the halves itself
                r32, d32, r16, d16 = ops.stem_s2d(rgb, depth, self.stem_packed, sr, br, sdp, bd, want_f32=learned,
                                                  split=True, bn_host=self.stem_bn_host)
                self.launches += 2
            elif self.se is None:
                r32, d32, _, _ = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=True)
                r16, d16 = ops.split_from_f32(r32), ops.split_from_f32(d32)
                self.launches += 4
            else:                                  # SE-add: squeeze pass, excite, SE-scaled stem (see below), then split
                part, inv_area = ops.stem_squeeze(rgb, depth, wr, sr, br, wd, sdp, bd)
                sig_r = ops.se_mlp(part, inv_area, *self.se[0]["se_rgb"], c_off=0, c=64)
                sig_d = ops.se_mlp(part, inv_area, *self.se[0]["se_depth"], c_off=64, c=64)
                r32, d32, _, _ = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=True, se_rgb=sig_r, se_depth=sig_d)
                r16, d16 = ops.split_from_f32(r32), ops.split_from_f32(d32)
                keep += [part, sig_r, sig_d]
                self.launches += 7
        elif self.stem_packed is not None:
            r32, d32, r16, d16 = ops.stem_s2d(rgb, depth, self.stem_packed, sr, br, sdp, bd, want_f32=learned,
                                              bn_host=self.stem_bn_host)
            self.launches += 2
        elif self.se is None:
            r32, d32, r16, d16 = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=learned)
            self.launches += 1
        else:
            # se_layer0 needs the global average of both FULL stem maps before they can be fused:
            # squeeze pass (channel sums only), excite (tiny MLP), then the fused stem with the scales
            part, inv_area = ops.stem_squeeze(rgb, depth, wr, sr, br, wd, sdp, bd)
            sig_r = ops.se_mlp(part, inv_area, *self.se[0]["se_rgb"], c_off=0, c=64)
            sig_d = ops.se_mlp(part, inv_area, *self.se[0]["se_depth"], c_off=64, c=64)
            r32, d32, r16, d16 = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=learned, se_rgb=sig_r,
                                          se_depth=sig_d)
            keep += [part, sig_r, sig_d]
            self.launches += 4
        if weight is not None:
            weight = weight.to(self.dev, torch.float32).contiguous()
        elif baseline:                                     # :264-266
            weight = torch.zeros(b, 5, device=self.dev)
            weight[:, 4] = 1
        elif ini_stage:                                    # :267-270, global CPU generator like the reference
            idx = torch.randint(0, 5, (b,))
            weight = torch.zeros(b, 5)
            weight[range(b), idx] = 1
            weight = weight.to(self.dev)
        # The learned gate (GlobalGate + DiffSoftmax + plan, ~0.1 ms of small kernels) is only needed by the depth
        # encoder and by the gated convolution that ENDS RGB stage 1: it runs at the head of the side stream, under
        # the first RGB convolutions, instead of in front of both encoders.
        gate_on_side = learned and not self.use_programs
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side if gate_on_side else main):
            if learned:
                # GlobalGate + DiffSoftmax + plan: two convolutions and one decision kernel
                gw = self.gate
                weight, plan, _ = ops.global_gate_decide(r32, d32, gw["w1"], gw["s1"], gw["b1"], gw["w2"], gw["s2"],
                                                         gw["b2"], gw["wfc"], temp, hard_gate, hist=hist)
                self.launches += 3
            else:
                plan = ops.gate_plan(weight, hist=hist)
                self.launches += 1
        keep += [r32, d32, r16, d16, weight, plan]

        self.programs = []
        skips = None
        if self.use_programs:
            c4 = self.stage_channels[3]
            cat = torch.empty(b, h // 32, w // 32, c4 + self.ppm_c, dtype=torch.bfloat16, device=self.dev)
            fused, skips = self._encoder_program(r16, d16, plan, cat, keep)
        else:
            # ---- depth encoder on the side stream, in slot order, prefix-counted
            if not gate_on_side:
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
            done = [torch.cuda.Event() for _ in range(4)]
            depth_out = []
            split = 1 if self.use_merge else 4          # stages run in the two-stream form
            with torch.cuda.stream(side):
                d = d16
                for s in range(split):
                    cnt = plan.count[s:s + 1]
                    for bi, blk in enumerate(self.stages["encoder_depth"][s]):
                        d = self._block(d, blk, keep, count=cnt, in_map=plan.perm if (s == 0 and bi == 0) else None,
                                        count_settled=not (s == 0 and bi == 0))
                    depth_out.append(d)
                    done[s].record(side)

            # the skip-connection convolutions (model.py:295-297) only feed the decoder: each one is launched on a
            # third stream the moment its stage output exists, and the decoder waits for it where it adds the skip
            skips = []
            skip_done = []

            def emit_skip(s, x):
                if s >= 3:
                    return
                if not self.dec_fusion:               # no skip connections into the decoder
                    skips.append(None)
                    skip_done.append(None)
                    return
                if self.skips[s] is None:
                    skips.append(x)
                    skip_done.append(None)
                    return
                ready = torch.cuda.Event()
                ready.record(main)
                self.side2.wait_event(ready)
                with torch.cuda.stream(self.side2):
                    y = self.skips[s](x)
                    ev = torch.cuda.Event()
                    ev.record(self.side2)
                skips.append(y)
                skip_done.append(ev)
                self.launches += 1

            # ---- RGB encoder on the main stream; the last conv of each stage adds g_s * depth_s
            r = r16
            fused = []
            cat = None
            for s in range(split):
                blocks = self.stages["encoder_rgb"][s]
                for bi, blk in enumerate(blocks):
                    if bi < len(blocks) - 1:
                        r = self._block(r, blk, keep)
                        continue
                    if s == 3:
                        # stage-4 output lands directly in the pyramid-pooling concat buffer
                        c4 = self.stage_channels[3]
                        cat = torch.empty(b, h // 32, w // 32, (c4 + self.ppm_c) * (2 if self.split else 1),
                                          dtype=torch.bfloat16, device=self.dev)
                    if self.se is None:
                        last_kw = dict(gated=depth_out[s], gate=plan.g[s], gated_slot=plan.slot)
                        if s == 3:
                            last_kw.update(out=cat, out_c_off=0)
                        r = self._block(r, blk, keep, before_last=lambda s=s: main.wait_event(done[s]), last_kw=last_kw)
                    else:
                        # SE-add: both stage outputs must be complete before they can be squeezed
                        r = self._block(r, blk, keep)
                        main.wait_event(done[s])
                        r = self._se_fuse(s, r, depth_out[s], plan, keep, cat if s == 3 else None)
                fused.append(r)
                emit_skip(s, r)
            for s in range(split, 4):                   # merged launches (the stage-1 join ordered main after side)
                if s == 3:
                    c4 = self.stage_channels[3]
                    cat = torch.empty(b, h // 32, w // 32, (c4 + self.ppm_c) * (2 if self.split else 1),
                                      dtype=torch.bfloat16, device=self.dev)
                r, d = self._merged_stage(s, r, d, plan, keep, cat)
                fused.append(r)
                emit_skip(s, r)

        # ---- context module, decoder (model.py:295-308, context_modules.py:69-87)
        if self.use_programs:
            skip_done = [None, None, None]
        out = self._decode(cat, skips, skip_done, keep, main, out, labels, want_logits)
        # all side-stream work was joined by the stage-4 wait; temporaries may now be released
        del keep
        return out, weight

    # ------------------------------------------------------------------ local gates (SkipESANet)
    def _local_gate_weight(self, i: int, gap_r: Tensor, gap_d: Tensor, alive: Optional[Tensor], prev: Optional[Tensor],
                           *, temp: float, hard: bool, random_policy: bool) -> Tensor:
        """SqueezeAndExciteReweigh.forward (rgb_depth_fusion.py:35-65) on the global average pools of the two
        streams: score = sigmoid(mean(x * SE(x))) only needs gap(x); Gumbel-softmax with the noise drawn like
        F.gumbel_softmax draws it; chained with the previous site's weight.  -> [B, 2] fp32."""
        from .local_gate import gumbel_softmax
        b = gap_r.shape[0]
        if random_policy:
            b0 = torch.randint(0, 2, (b,))                       # global CPU generator, like the reference (:38)
            w = torch.stack([b0, 1 - b0], dim=1).to(self.dev, torch.float32)
        else:
            if alive is not None:                                # closed samples have no depth features: any finite
                gap_d = torch.where(alive.view(-1, 1), gap_d, torch.zeros_like(gap_d))   # value (the chain zeroes w)
            w1, b1, w2, b2 = self.local_gates[i]
            gap = torch.cat([gap_r, gap_d], dim=1)               # gap(cat(rgb, depth)), never materialised
            se = torch.sigmoid(torch.relu(gap @ w1.t() + b1) @ w2.t() + b2)
            score = torch.sigmoid((gap * se).mean(dim=1))
            w = gumbel_softmax(torch.stack([score, 1 - score], dim=1) / temp, hard)
        if prev is not None:
            b1_ = w[:, 1] * prev
            w = torch.stack([1 - b1_, b1_], dim=1)
        return w

    @torch.no_grad()
    def forward_local(self, rgb: Tensor, depth: Tensor, *, block_rule: Sequence[int], temp: float = 1.0,
                      hard: bool = True, random_policy: bool = False, ini_stage: bool = False,
                      out: Optional[Tensor] = None):
        """``SkipESANet.forward`` in eval mode (model_skip_mod.py:246-322) on the CUDA kernels, with REAL skipping:
        site s decides before stage s + 1 runs, so the depth encoder's stage s + 1 only processes the samples that
        can still use depth features -- the blend of this site (rule 1, or rule 2 with a non-zero weight) or, through
        the chain w_t[1] *= w_{t-1}[1], a later one.  Those samples are compacted to the front of the stage's
        tensors (slot order; ``count`` / ``in_map`` / ``gated_slot`` of the conv kernels), recomputed per stage from
        the gate weights on the device: no host synchronisation, CUDA-graph capturable.
        -> (logits [B,classes,H,W] fp32, [w_0 .. w_3] gate weights [B,2] fp32, [count_1 .. count_4] device int32)"""
        with torch.cuda.device(self.dev):
            return self._forward_local(rgb, depth, block_rule=list(block_rule), temp=temp, hard=hard,
                                       random_policy=random_policy, ini_stage=ini_stage, out=out)

    def _forward_local(self, rgb, depth, *, block_rule, temp, hard, random_policy, ini_stage, out):
        if self.cfg.gate != "local":
            raise _lib.DynmmError("forward_local needs an engine built with gate='local'")
        rgb = rgb.float().contiguous()
        depth = depth.float().contiguous()
        b, _, h, w = rgb.shape
        if h % 32 or w % 32:
            raise _lib.DynmmError(f"input size {h}x{w} must be a multiple of 32 (five stride-2 stages)")
        self.launches = 0
        keep: list = []
        main = torch.cuda.current_stream()
        side = self.side
        dev = self.dev
        wr, sr, br = self.stem["encoder_rgb"]
        wd, sdp, bd = self.stem["encoder_depth"]
        dynamic = [r == 2 for r in block_rule]
        gate_kw = dict(temp=temp, hard=hard, random_policy=random_policy)
        # ---- stem; gate 0 looks at the UNPOOLED, unfused stem maps (model_skip_mod.py:253): squeeze pass
        weights: List[Tensor] = []
        if dynamic[0] and not random_policy:
            part, inv_area = ops.stem_squeeze(rgb, depth, wr, sr, br, wd, sdp, bd)
            gap = part.sum(dim=1) * inv_area                          # [B, 128] = [rgb 64 | depth 64]
            weights.append(self._local_gate_weight(0, gap[:, :64], gap[:, 64:], None, None, **gate_kw))
            keep.append(part)
            self.launches += 1
        else:
            # the reference evaluates every gate (its Gumbel / randint draws advance the generator) even where the
            # rule ignores the result: keep the draw order
            z = torch.zeros(b, 64, device=dev)
            weights.append(self._local_gate_weight(0, z, z, None, None, **gate_kw))
        if self.stem_packed is not None:
            _, _, r16, d16 = ops.stem_s2d(rgb, depth, self.stem_packed, sr, br, sdp, bd, want_f32=False,
                                          split=self.split, bn_host=self.stem_bn_host)
            self.launches += 2
        elif self.split:
            r32, d32, _, _ = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=True)
            r16, d16 = ops.split_from_f32(r32), ops.split_from_f32(d32)
            self.launches += 3
        else:
            _, _, r16, d16 = ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd, want_f32=False)
            self.launches += 1
        keep += [r16, d16]

        ones = torch.ones(b, dtype=torch.bool, device=dev)
        alive = ones                     # chain still open: every dynamic site so far kept a non-zero depth weight
        prev_w: Optional[Tensor] = None  # w_{s-1}[:, 1] of the last dynamic site (None with ini_stage)
        prev_slot: Optional[Tensor] = None   # sample -> slot in the previous depth tensor
        r, d = r16, d16
        fused, counts = [], []
        skips, skip_done = [], []
        cat = None
        arange = torch.arange(b, device=dev, dtype=torch.int32)
        for s in range(4):
            rule = block_rule[s]
            wgt = weights[s]
            if rule == 2:
                g = wgt[:, 1].contiguous()
                alive = alive & (g != 0) if not ini_stage else alive
            elif rule == 1:
                g = torch.ones(b, device=dev)
            else:
                g = torch.zeros(b, device=dev)
            # who needs depth stage s + 1: this site's blend, a later static add, or a later dynamic site whose chain
            # is still open (without chaining -- ini_stage -- every later dynamic site may still open)
            later_add = any(rr == 1 for rr in block_rule[s + 1:])
            later_dyn = any(rr == 2 for rr in block_rule[s + 1:])
            need = (g != 0)
            if later_add:
                need = ones
            elif later_dyn:
                need = need | alive
            if s > 0:
                need = need & (prev_slot >= 0)                        # features of a closed chain no longer exist
            # slot order of this stage: needed samples first (stable), the rest behind `count`
            order = torch.argsort((~need).to(torch.int8), stable=True).to(torch.int32)
            count = need.sum().to(torch.int32).reshape(1)
            slot = torch.empty(b, dtype=torch.int32, device=dev)
            slot[order.long()] = arange
            slot = torch.where(need, slot, torch.full_like(slot, -1))
            in_map = (order if prev_slot is None else prev_slot[order.long()].clamp_min(0)).to(torch.int32).contiguous()
            gated_slot = slot.clamp_min(0).contiguous()
            counts.append(count)
            keep += [order, count, slot, in_map, gated_slot, g, need]
            # ---- depth stage on the side stream
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
            done = torch.cuda.Event()
            with torch.cuda.stream(side):
                for bi, blk in enumerate(self.stages["encoder_depth"][s]):
                    d = self._block(d, blk, keep, count=count, in_map=in_map if bi == 0 else None)
                done.record(side)
            # ---- RGB stage on the main stream; rule != 0: the last conv adds g * depth
            blocks = self.stages["encoder_rgb"][s]
            for bi, blk in enumerate(blocks):
                if bi < len(blocks) - 1:
                    r = self._block(r, blk, keep)
                    continue
                last_kw = {}
                if s == 3:
                    c4 = self.stage_channels[3]
                    cat = torch.empty(b, h // 32, w // 32, (c4 + self.ppm_c) * (2 if self.split else 1),
                                      dtype=torch.bfloat16, device=dev)
                    last_kw.update(out=cat, out_c_off=0)
                if rule != 0:
                    last_kw.update(gated=d, gate=g, gated_slot=gated_slot)
                    r = self._block(r, blk, keep, before_last=lambda: main.wait_event(done), last_kw=last_kw)
                else:
                    r = self._block(r, blk, keep, last_kw=last_kw or None)
                    main.wait_event(done)                              # join the side stream (gate features below)
            fused.append(r)
            if s < 3:
                # skip connection conv of this stage (decoder input)
                if not self.dec_fusion:
                    skips.append(None)
                elif self.skips[s] is not None:
                    skips.append(self.skips[s](r))
                    self.launches += 1
                else:
                    skips.append(r)
                skip_done.append(None)
                # gate s + 1 looks at the stage outputs BEFORE the blend (:260, :278, :296):
                # gap(rgb) = gap(fuse) - g * gap(depth) (the blend is linear)
                c = r.shape[3] // (2 if self.split else 1)
                if dynamic[s + 1] and not random_policy:
                    inv_area = 1.0 / (r.shape[1] * r.shape[2])
                    pf = ops.gap_partial(r, c=c, split=self.split)
                    pd = ops.gap_partial(d, c=c, count=count, split=self.split)
                    gap_f = pf.sum(dim=1) * inv_area
                    gap_d_slots = pd.sum(dim=1) * inv_area
                    has = slot >= 0
                    gap_d = torch.where(has.view(-1, 1), gap_d_slots[gated_slot.long()], torch.zeros_like(gap_f))
                    gap_r = gap_f - g.view(-1, 1) * gap_d
                    keep += [pf, pd]
                    self.launches += 2
                    nxt = self._local_gate_weight(s + 1, gap_r, gap_d, has, None, **gate_kw)
                else:
                    z = torch.zeros(b, c, device=dev)
                    nxt = self._local_gate_weight(s + 1, z, z, None, None, **gate_kw)
                if rule == 2 and not ini_stage:
                    prev_w = weights[s][:, 1]
                if prev_w is not None:                                  # chain (rgb_depth_fusion.py:60-63)
                    b1_ = nxt[:, 1] * prev_w
                    nxt = torch.stack([1 - b1_, b1_], dim=1)
                weights.append(nxt)
            prev_slot = slot
        out = self._decode(cat, skips, skip_done, keep, main, out, None, True)
        del keep
        return out, weights, counts

    def _decode(self, cat: Tensor, skips: list, skip_done: list, keep: list, main, out, labels, want_logits):
        """Pyramid pooling, the three decoder modules and the learned final up-samplings (model.py:295-308,
        context_modules.py:69-87) on the stage-4 output in ``cat[..., :c4]`` and the three skip tensors."""
        c4 = self.stage_channels[3]
        off = c4
        if self.use_programs:
            pooled = [ops.adaptive_avgpool(cat, bins, c=c4) for bins in self.ppm_bins]
            with ops.ConvProgram() as prog:          # the pyramid branches are independent: one phase
                ys = [self.ppm[i](pooled[i]) for i in range(len(self.ppm_bins))]
            self.programs.append(prog)
            for y in ys:
                (ops.bilinear_resize_into if self.ctx_bilinear else ops.nearest_resize_into)(y, cat, off)
                off += y.shape[3]
            cat._dynmm_flags = None             # other kernels wrote into the buffer since the conv published its flags
            keep += pooled + ys
            self.launches += 3 * len(self.ppm_bins)
        else:
            sp = self.split
            for i, bins in enumerate(self.ppm_bins):
                pooled = ops.adaptive_avgpool(cat, bins, c=c4, split=sp)
                y = self.ppm[i](pooled)
                (ops.bilinear_resize_into if self.ctx_bilinear else ops.nearest_resize_into)(y, cat, off, split=sp)
                off += y.shape[3] // (2 if sp else 1)
                keep += [pooled, y]
                self.launches += 3
            cat._dynmm_flags = None             # other kernels wrote into the buffer since the conv published its flags
        keep += [cat] + [t for t in skips if t is not None]
        if self.use_programs:
            x = cat
            for i, skip in enumerate((skips[2], skips[1], skips[0])):
                m = self.dec[i]
                items = ([self.ppm_final] if i == 0 else []) + [m["conv3x3"]] + list(m["blocks"])
                x = self._sequence_program(x, items, keep)
                x = ops.upsample2x_dw3x3(x, m["up_w"], m["up_b"], skip, replicate=self.up_replicate)
                self.launches += 1
                keep.append(x)
        else:
            x = self.ppm_final(cat)
            self.launches += 1
            keep.append(x)
            for i, skip in enumerate((skips[2], skips[1], skips[0])):
                m = self.dec[i]
                x = m["conv3x3"](x)
                self.launches += 1
                keep.append(x)
                for blk in m["blocks"]:
                    x = self._block(x, blk, keep)
                if skip_done[2 - i] is not None:
                    main.wait_event(skip_done[2 - i])
                x = ops.upsample2x_dw3x3(x, m["up_w"], m["up_b"], skip, split=self.split, replicate=self.up_replicate)
                self.launches += 1
                keep.append(x)
        x = self.conv_out(x)
        keep.append(x)
        x = ops.upsample2x_dw3x3(x, self.up[0][0], self.up[0][1], split=self.split, replicate=self.up_replicate)
        keep.append(x)
        out = ops.upsample2x_dw3x3(x, self.up[1][0], self.up[1][1], to_nchw_f32=True, out=out, labels=labels,
                                   want_logits=want_logits, split=self.split, replicate=self.up_replicate,
                                   c_valid=self.n_classes)
        self.launches += 3
        return out
