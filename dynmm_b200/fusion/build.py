"""``build_model(args, n_classes)`` -- drop-in for FusionDynMM/src/build_model.py:18-218
restricted to the dynamic models this package provides: ``--dynamic --global-gate`` (``SkipGateESANet``, the
accelerated path) and ``--dynamic`` alone (``SkipESANet``, the local-gate variant, build_model.py:76-95)."""
from __future__ import annotations

import warnings

import torch
from torch import nn

from .local_gate import SkipESANet
from .modules import SkipGateESANet


def build_model(args, n_classes):
    if not getattr(args, "dynamic", False):
        raise NotImplementedError("dynmm_b200 provides the dynamic models (--dynamic [--global-gate]); the static "
                                  "ESANet / one-modality variants are outside the gated path")
    if getattr(args, "pretrained_on_imagenet", False) and not getattr(args, "last_ckpt", "") and \
            getattr(args, "pretrained_scenenet", "") == "":
        warnings.warn("ImageNet checkpoints are not reachable offline: building with random init")
    if "decreasing" in args.decoder_channels_mode:
        channels_decoder = [512, 256, 128]            # build_model.py:27-32
    else:
        channels_decoder = [args.channels_decoder] * 3
    nr = args.nr_decoder_blocks
    nr = [nr] * 3 if isinstance(nr, int) else (list(nr) * 3 if len(nr) == 1 else list(nr))
    assert len(nr) == 3
    block_rule = [int(s) for s in args.block_rule]
    assert len(block_rule) == 4
    if args.encoder_depth in (None, "None"):
        args.encoder_depth = args.encoder
    cls = SkipGateESANet if getattr(args, "global_gate", False) else SkipESANet      # build_model.py:54-95
    model = cls(
        height=args.height, width=args.width, num_classes=n_classes, pretrained_on_imagenet=False,
        pretrained_dir=getattr(args, "pretrained_dir", None), encoder_rgb=args.encoder,
        encoder_depth=args.encoder_depth, encoder_block=args.encoder_block, activation=args.activation,
        encoder_decoder_fusion=args.encoder_decoder_fusion, context_module=args.context_module,
        nr_decoder_blocks=nr, channels_decoder=channels_decoder,
        fuse_depth_in_rgb_encoder=args.fuse_depth_in_rgb_encoder, upsampling=args.upsampling, temp=args.temp,
        block_rule=block_rule)
    pretrained_on_imagenet = bool(getattr(args, "pretrained_on_imagenet", False)) and \
        not getattr(args, "last_ckpt", "") and getattr(args, "pretrained_scenenet", "") == ""      # build_model.py:19-23
    device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")      # build_model.py:143-149
    model.to(device)                                  # BEFORE the init below, as in the reference (RNG stream)
    if getattr(args, "he_init", False):
        he_init(model, n_classes, pretrained_on_imagenet)
    scenenet = getattr(args, "pretrained_scenenet", "")
    if scenenet != "":                                                            # build_model.py:180-206
        load_scenenet(model, scenenet, args.context_module)
    if getattr(args, "finetune", None):                                           # build_model.py:208-211
        ckpt = torch.load(args.finetune, map_location="cpu", weights_only=False)
        model.load_state_dict(ckpt["state_dict"], strict=False)
    return model, device


def he_init(model: nn.Module, n_classes: int, pretrained_on_imagenet: bool) -> None:
    """The reference's ``--he_init`` loop (build_model.py:152-178), as written: walk ``model.children()`` (the
    encoders are skipped only when they carry ImageNet weights), Kaiming-normal(fan_out, relu) every Conv2d / Conv1d /
    Linear EXCEPT output layers (``out_channels == n_classes``: conv_out, side outputs), layers followed by a Sigmoid
    (the second 1x1 of an SE block) and depthwise convolutions (learned upsampling); biases are never touched;
    BatchNorm / GroupNorm get weight 1, bias 0.  The gate convolutions and (without ImageNet weights) the encoders
    are re-initialised too.

    One deviation: the reference indexes ``module_list[i + 1]`` unguarded, which raises IndexError when the LAST
    module of the model is a convolution -- the case for ``SkipGateESANet`` (``gate_layer.fc``, :385); here "no next
    module" simply means "not followed by a Sigmoid"."""
    from .modules import ResNet
    module_list = []
    for c in model.children():
        if pretrained_on_imagenet and isinstance(c, ResNet):
            continue                                  # already initialised
        module_list.extend(c.modules())
    for i, m in enumerate(module_list):
        if isinstance(m, (nn.Conv2d, nn.Conv1d, nn.Linear)):
            out_channels = m.out_features if isinstance(m, nn.Linear) else m.out_channels
            nxt = module_list[i + 1] if i + 1 < len(module_list) else None
            groups = getattr(m, "groups", 1)
            in_channels = m.in_features if isinstance(m, nn.Linear) else m.in_channels
            if out_channels == n_classes or isinstance(nxt, nn.Sigmoid) or groups == in_channels:
                continue
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def load_scenenet(model: nn.Module, path: str, context_module: str) -> None:
    """``--pretrained_scenenet`` (build_model.py:180-206): update the model's state with the checkpoint's, minus the
    (side) outputs, the learned final upsamplings and -- for context modules other than ppm / appm -- the pyramid
    features; strict load of the merged dict."""
    checkpoint = torch.load(path, map_location="cpu", weights_only=False)
    weights = dict(checkpoint["state_dict"])
    ignore = [k for k in weights if "out" in k or "decoder.upsample1" in k or "decoder.upsample2" in k]
    if context_module not in ("ppm", "appm"):
        ignore.extend(k for k in weights if "context_module.features" in k)
    for k in set(ignore):
        weights.pop(k)
    merged = model.state_dict()
    merged.update(weights)
    model.load_state_dict(merged)
