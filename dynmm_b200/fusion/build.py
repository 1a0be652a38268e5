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
    device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")      # build_model.py:143-149
    if getattr(args, "he_init", False):                                           # build_model.py:152-178
        for name, m in model.named_modules():
            if "encoder" in name or "gate" in name:
                continue
            if isinstance(m, nn.Conv2d) and m.groups == 1:
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
    if getattr(args, "finetune", None):                                           # build_model.py:208-211
        ckpt = torch.load(args.finetune, map_location="cpu", weights_only=False)
        model.load_state_dict(ckpt["state_dict"], strict=False)
    model.to(device)
    return model, device
