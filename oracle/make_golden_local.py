"""Generate ``tests/golden/local_gate_*.npz`` by running the REFERENCE's local-gate model
(``FusionDynMM/src/models/model_skip_mod.py`` ``SkipESANet``) -- SURVEY.md section 8f-4.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference`` does not exist on the GPU box):
``python -m oracle.make_golden_local``.  Nothing is copied from the reference: its module is imported from where it
lies, gets the seeded state of ``seeded_state`` below loaded with ``strict=True`` (which pins the state_dict key names
and shapes of the drop-in class) and its outputs are stored.  Every forward is preceded by ``torch.manual_seed`` so the
Gumbel noise (``F.gumbel_softmax``, rgb_depth_fusion.py:50,56) and the random policy (``torch.randint``, :38) are
reproducible by anything that draws in the same order from the same generator.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference/FusionDynMM"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (constructor kwargs, state seed, batch)
CASES = {
    "local_gate_r18_basic_64x64": (dict(height=64, width=64, num_classes=37, encoder_rgb="resnet18",
                                        encoder_depth="resnet18", encoder_block="BasicBlock",
                                        fuse_depth_in_rgb_encoder="SE-add", upsampling="bilinear"), 5, 3),
    "local_gate_r34_nbt1d_64x96": (dict(height=64, width=96, num_classes=40, encoder_rgb="resnet34",
                                        encoder_depth="resnet34", encoder_block="NonBottleneck1D",
                                        nr_decoder_blocks=[3, 3, 3], fuse_depth_in_rgb_encoder="add",
                                        upsampling="learned-3x3-zeropad"), 6, 2),
}
# (tag, block_rule, attributes, test flag, forward seed)
MODES = [
    ("test_hard", [2, 2, 2, 2], dict(), True, 11),
    ("soft", [2, 2, 2, 2], dict(), False, 12),
    ("hard", [2, 2, 2, 2], dict(hard_gate=True), False, 13),
    ("ini", [2, 2, 2, 2], dict(hard_gate=True, ini_stage=True), False, 14),
    ("random", [2, 2, 2, 2], dict(random_policy=True), True, 15),
    ("mixed1122", [1, 1, 2, 2], dict(hard_gate=True), True, 16),
    ("mixed0120", [0, 1, 2, 0], dict(), True, 17),
    ("static1111", [1, 1, 1, 1], dict(), True, 18),
]


def seeded_state(state_dict, seed: int):
    """A reproducible state for any ESANet-style state_dict: every tensor is redrawn from one seeded generator in sorted
    key order -- conv / linear weights ~ N(0, 2/fan_in), biases and BN shifts ~ N(0, 0.1), BN scales in [0.6, 1.0],
    running means ~ N(0, 0.1), running variances in [0.5, 1.5]."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith("weight"):          # BatchNorm scale
            out[k] = 0.6 + 0.4 * torch.rand(v.shape, generator=g)
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
    return out


def sample_inputs(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    rgb, depth = torch.randn(b, 3, h, w, generator=g), torch.randn(b, 1, h, w, generator=g)
    gain = 0.25 + 1.5 * torch.rand(b, 2, generator=g)
    return rgb * gain[:, 0].view(-1, 1, 1, 1), depth * gain[:, 1].view(-1, 1, 1, 1) + 0.3


def apply_mode(model, rule, attrs):
    model.block_rule = list(rule)
    model.hard_gate = model.ini_stage = model.random_policy = False
    for k, v in attrs.items():
        setattr(model, k, v)


def main():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present; golden vectors can only be regenerated in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from src.models.model_skip_mod import SkipESANet            # the reference implementation
    os.makedirs(OUT, exist_ok=True)
    for name, (kw, seed, b) in CASES.items():
        model = SkipESANet(pretrained_on_imagenet=False, **kw)
        sd = seeded_state(model.state_dict(), seed)
        model.load_state_dict(sd, strict=True)
        rgb, depth = sample_inputs(seed + 100, b, kw["height"], kw["width"])
        res = {"keys": np.array(sorted(sd.keys())), "seed": np.array(seed), "batch": np.array(b)}
        model.eval()
        with torch.no_grad():
            for tag, rule, attrs, test, fseed in MODES:
                apply_mode(model, rule, attrs)
                model.start_weight()
                torch.manual_seed(fseed)
                out = model(rgb, depth, test)
                for i in range(4):
                    res[f"{tag}_weight{i}"] = model.weight_list[i].numpy()
                model.end_weight()
                res[f"{tag}_out"] = out[:, :, ::4, ::4].contiguous().numpy()
                res[f"{tag}_abssum"] = np.float64(out.double().abs().sum().item())
                print(name, tag, [res[f"{tag}_weight{i}"][:, 1].round(3).tolist() for i in range(4)])
        # training mode: 4 scales, batch-statistics BN, soft Gumbel gates
        model.train()
        apply_mode(model, [2, 2, 2, 2], {})
        torch.manual_seed(21)
        with torch.no_grad():
            outs = model(rgb, depth)
        for i, o in enumerate(outs):
            res[f"train_out{i}_shape"] = np.array(o.shape)
            res[f"train_out{i}_abssum"] = np.float64(o.double().abs().sum().item())
        res["train_out0"] = outs[0][:, :, ::4, ::4].contiguous().numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)


if __name__ == "__main__":
    main()
