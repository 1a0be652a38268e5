"""Generate ``tests/golden/*.npz`` by running the REFERENCE itself.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python -m oracle.make_golden``.  The reference modules are imported
from where they lie -- nothing is copied.  Weights come from
``oracle.fusion_oracle.make_state_dict`` (seeded, reproducible anywhere) and
are loaded into the reference ``SkipGateESANet`` with ``strict=True``, which
also pins the state_dict key names and shapes.

Stored per case: inputs' seed, the gate weights, the FLOP loss, gate logits
(captured with a forward hook on ``gate_layer.fc``), a strided sample of the
logits plus global sums (keeps fixtures small), and the fused stage outputs'
per-sample means.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference/FusionDynMM"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    """The reference hard-codes ``.cuda()`` (model_skip_mod_globalgate.py:218-223,
    249,265,268); on this CUDA-less host make it an identity."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present; golden vectors can only be regenerated in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    warnings.filterwarnings("ignore")
    from src.models import model_skip_mod_globalgate as m
    return m


def sample_inputs(seed, b, h, w):
    """N(0,1) images with a seeded per-sample gain/offset so that an untrained
    gate does not put the whole batch on one branch."""
    g = torch.Generator().manual_seed(seed)
    rgb, depth = torch.randn(b, 3, h, w, generator=g), torch.randn(b, 1, h, w, generator=g)
    gain = 0.25 + 1.5 * torch.rand(b, 2, generator=g)
    off = torch.randn(b, 2, generator=g)
    rgb = rgb * gain[:, 0].view(-1, 1, 1, 1) + off[:, 0].view(-1, 1, 1, 1)
    depth = depth * gain[:, 1].view(-1, 1, 1, 1) + off[:, 1].view(-1, 1, 1, 1)
    return rgb, depth


def summarize(prefix, res, t):
    t = t.detach()
    res[prefix + "_sample"] = t[:, :, ::4, ::4].contiguous().numpy()
    res[prefix + "_sum"] = np.float64(t.double().sum().item())
    res[prefix + "_abssum"] = np.float64(t.double().abs().sum().item())
    res[prefix + "_shape"] = np.array(t.shape)


def build_reference_model(m, cfg, sd):
    model = m.SkipGateESANet(height=cfg.height, width=cfg.width, num_classes=cfg.num_classes,
                             encoder_rgb=cfg.encoder, encoder_depth=cfg.encoder,
                             encoder_block=cfg.encoder_block, channels_decoder=list(cfg.channels_decoder),
                             nr_decoder_blocks=list(cfg.nr_decoder_blocks),
                             fuse_depth_in_rgb_encoder=cfg.fuse_depth_in_rgb_encoder,
                             context_module=cfg.context_module, upsampling=cfg.upsampling,
                             activation=cfg.activation)
    model.load_state_dict(sd, strict=True)
    return model


def run_fusion_case(m, name, cfg, seed, b, gate_scale=40.0):
    from oracle import fusion_oracle as fo
    sd = fo.make_state_dict(cfg, seed, gate_scale)
    model = build_reference_model(m, cfg, sd)
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    res = {"seed": np.array(seed), "batch": np.array(b), "gate_scale": np.array(gate_scale),
           "keys": np.array(sorted(model.state_dict().keys()))}
    cap = {}
    hook = model.gate_layer.fc.register_forward_hook(lambda mod, i, o: cap.__setitem__("logits", o.detach().flatten(1)))
    fuse_cap = []

    model.eval()
    with torch.no_grad():
        # learned gate, soft and hard, two temperatures
        for tag, temp, hard in (("soft_t1", 1.0, False), ("hard_t1", 1.0, True), ("soft_t01", 0.1, False)):
            model.temp, model.hard_gate, model.baseline, model.ini_stage = temp, hard, False, False
            out, w = model(rgb, depth, True, True)
            summarize(f"{tag}_out", res, out)
            res[f"{tag}_weight"] = w.numpy()
            res[f"{tag}_logits"] = cap["logits"].numpy()
        # baseline = static ESANet
        model.baseline = True
        out, w = model(rgb, depth, True, True)
        summarize("baseline_out", res, out)
        res["baseline_weight"] = w.numpy()
        model.baseline = False
        # forced branches through the reference's own ini_stage path (CPU RNG, :267-270)
        model.ini_stage = True
        torch.manual_seed(1234)
        out, w = model(rgb, depth, True, True)
        summarize("ini_out", res, out)
        res["ini_weight"] = w.numpy()
        model.ini_stage = False
        # every one-hot branch, all samples
        orig = model.gate_layer.forward
        for k in range(5):
            model.gate_layer.forward = lambda r, d, t=1.0, h=False, k=k: torch.eye(5)[torch.full((r.shape[0],), k)]
            out, w = model(rgb, depth, True, True)
            summarize(f"branch{k}_out", res, out)
        model.gate_layer.forward = orig
    # training mode: 4 scales + FLOP loss, batch-stat BN
    model.train()
    model.temp, model.hard_gate = 1.0, False
    with torch.no_grad():
        outs, loss = model(rgb, depth)
    for i, o in enumerate(outs):
        t = o.detach()
        res[f"train_out{i}_sum"] = np.float64(t.double().sum().item())
        res[f"train_out{i}_abssum"] = np.float64(t.double().abs().sum().item())
        res[f"train_out{i}_shape"] = np.array(t.shape)
        res[f"train_out{i}_sample"] = t[:, :, ::4, ::4].contiguous().numpy() if i == 0 else t.numpy()
    res["train_loss"] = np.float64(loss.item())
    hook.remove()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, "hard branches:", res["hard_t1_weight"].argmax(1), "ini:", res["ini_weight"].argmax(1))


def run_diffsoftmax(m):
    g = torch.Generator().manual_seed(7)
    res = {}
    cases = {"b16x5": torch.randn(16, 5, generator=g) * 3, "b128x2": torch.randn(128, 2, generator=g),
             "ties": torch.tensor([[1., 1., 0., 0., 0.], [0., 2., 2., 2., 0.], [5., 5., 5., 5., 5.],
                                   [-1., 0., 0.5, 0.5, 0.25]])}
    for name, logits in cases.items():
        res[name + "_logits"] = logits.numpy()
        up = torch.randn(logits.shape, generator=g)
        res[name + "_upstream"] = up.numpy()
        for tau in (1.0, 0.5, 1e-3):
            for hard in (False, True):
                x = logits.clone().requires_grad_(True)
                y = m.DiffSoftmax(x, tau=tau, hard=hard, dim=1)
                (y * up).sum().backward()
                tag = f"{name}_tau{tau}_{'hard' if hard else 'soft'}"
                res[tag + "_y"] = y.detach().numpy()
                res[tag + "_grad"] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "diffsoftmax.npz"), **res)
    print("diffsoftmax ok")


def main():
    from oracle import fusion_oracle as fo
    m = import_reference()
    os.makedirs(OUT, exist_ok=True)
    run_diffsoftmax(m)
    run_fusion_case(m, "fusion_r34_nbt1d_add_64x96", fo.FusionConfig(height=64, width=96), seed=0, b=4)
    run_fusion_case(m, "fusion_r34_nbt1d_seadd_64x64",
                    fo.FusionConfig(height=64, width=64, fuse_depth_in_rgb_encoder="SE-add"), seed=3, b=3)
    run_fusion_case(m, "fusion_r18_basic_add_64x64",
                    fo.FusionConfig(height=64, width=64, encoder="resnet18", encoder_block="BasicBlock"), seed=5, b=2)
    run_fusion_case(m, "fusion_r50_seadd_decr_64x64",
                    fo.FusionConfig(height=64, width=64, encoder="resnet50", encoder_block="BasicBlock",
                                    fuse_depth_in_rgb_encoder="SE-add", channels_decoder=(512, 256, 128)),
                    seed=9, b=2)


if __name__ == "__main__":
    main()
