"""Pin the CPU oracle against vectors produced by the REFERENCE itself
(oracle/make_golden.py, run where /root/reference exists)."""
import os

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle.make_golden import sample_inputs

CASES = {
    "fusion_r34_nbt1d_add_64x96": (fo.FusionConfig(height=64, width=96), 0, 4),
    "fusion_r34_nbt1d_seadd_64x64": (fo.FusionConfig(height=64, width=64, fuse_depth_in_rgb_encoder="SE-add"), 3, 3),
    "fusion_r18_basic_add_64x64": (fo.FusionConfig(height=64, width=64, encoder="resnet18",
                                                   encoder_block="BasicBlock"), 5, 2),
    "fusion_r50_seadd_decr_64x64": (fo.FusionConfig(height=64, width=64, encoder="resnet50", encoder_block="BasicBlock",
                                                    fuse_depth_in_rgb_encoder="SE-add",
                                                    channels_decoder=(512, 256, 128)), 9, 2),
}


def _close(t, gold, prefix, rtol=2e-4):
    sample = gold[prefix + "_sample"]
    scale = float(np.abs(sample).max())
    got = t[:, :, ::4, ::4].numpy()
    assert got.shape == sample.shape
    np.testing.assert_allclose(got, sample, rtol=rtol, atol=rtol * scale)
    assert abs(t.double().abs().sum().item() - gold[prefix + "_abssum"]) <= rtol * gold[prefix + "_abssum"]
    assert list(t.shape) == list(gold[prefix + "_shape"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_vectors(name, golden_dir):
    cfg, seed, b = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = fo.make_state_dict(cfg, seed, float(gold["gate_scale"]))
    assert sorted(sd.keys()) == list(gold["keys"])        # state_dict key names (strict load in eval.py:61)
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    with torch.no_grad():
        for tag, temp, hard in (("soft_t1", 1.0, False), ("hard_t1", 1.0, True), ("soft_t01", 0.1, False)):
            r = fo.forward(sd, cfg, rgb, depth, temp=temp, hard_gate=hard)
            np.testing.assert_allclose(r["gate_logits"].numpy(), gold[tag + "_logits"], rtol=1e-4, atol=1e-5)
            if hard:
                np.testing.assert_array_equal(r["weight"].numpy(), gold[tag + "_weight"])   # bit-exact decisions
            else:
                np.testing.assert_allclose(r["weight"].numpy(), gold[tag + "_weight"], rtol=1e-4, atol=1e-6)
            _close(r["out"], gold, tag + "_out")
        r = fo.forward(sd, cfg, rgb, depth, baseline=True)
        np.testing.assert_array_equal(r["weight"].numpy(), gold["baseline_weight"])
        _close(r["out"], gold, "baseline_out")
        torch.manual_seed(1234)
        r = fo.forward(sd, cfg, rgb, depth, ini_stage=True)
        np.testing.assert_array_equal(r["weight"].numpy(), gold["ini_weight"])
        _close(r["out"], gold, "ini_out")
        for k in range(5):
            w = torch.eye(5)[torch.full((b,), k)]
            r = fo.forward(sd, cfg, rgb, depth, weight=w)
            _close(r["out"], gold, f"branch{k}_out")
            # what the CUDA engine computes (depth stages really skipped) is the same function
            r2 = fo.forward(sd, cfg, rgb, depth, weight=w, skip_compute=True)
            _close(r2["out"], gold, f"branch{k}_out")
        r = fo.forward(sd, cfg, rgb, depth, temp=1.0, hard_gate=False, training=True)
        outs = r["out"]
        assert len(outs) == 4
        for i, o in enumerate(outs):
            assert list(o.shape) == list(gold[f"train_out{i}_shape"])
            ref = gold[f"train_out{i}_abssum"]
            assert abs(o.double().abs().sum().item() - ref) <= 5e-4 * ref
        assert abs(r["loss"].item() - float(gold["train_loss"])) <= 1e-5 * abs(float(gold["train_loss"])) + 1e-7


def test_diffsoftmax_matches_reference_vectors(golden_dir):
    gold = np.load(os.path.join(golden_dir, "diffsoftmax.npz"))
    for name in ("b16x5", "b128x2", "ties"):
        logits = torch.from_numpy(gold[name + "_logits"])
        up = torch.from_numpy(gold[name + "_upstream"])
        for tau in (1.0, 0.5, 1e-3):
            for hard in (False, True):
                tag = f"{name}_tau{tau}_{'hard' if hard else 'soft'}"
                x = logits.clone().requires_grad_(True)
                y = fo.diff_softmax(x, tau, hard, 1)
                (y * up).sum().backward()
                if hard:
                    np.testing.assert_array_equal(y.detach().numpy(), gold[tag + "_y"])
                else:
                    np.testing.assert_allclose(y.detach().numpy(), gold[tag + "_y"], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(x.grad.numpy(), gold[tag + "_grad"], rtol=1e-5, atol=1e-6)


def test_flop_tables_are_reproduced():
    """The reference's constant tables (model_skip_mod_globalgate.py:217-220) are
    its only known-answer vectors: conv-only MACs of the depth encoder,
    accumulated by branch, must land within thop's BN/activation overhead."""
    cfg = fo.FusionConfig()
    m = fo.conv_macs(cfg)
    cum = m["depth_stem"]
    table = fo.DEPTH_ENC_FLOP_R34
    assert abs(cum / 1e9 - table[0]) / table[0] < 0.05
    for s in range(4):
        cum += m[f"encoder_depth.stage{s + 1}"]
        assert abs(cum / 1e9 - table[s + 1]) / table[s + 1] < 0.02, (s, cum / 1e9, table[s + 1])
    # each 3x1 / 1x3 conv of a stage-1 block is 235.9 MMAC
    assert m["encoder_rgb.stage1"] == 3 * 4 * 64 * 64 * 3 * 120 * 160


def test_multiscale_cross_entropy_oracle_and_module_match_reference_vectors(golden_dir):
    """CrossEntropyLoss2d (utils.py:18-50).  The numpy oracle and the drop-in module against losses and
    logit gradients produced by the reference class (oracle/make_golden_loss.py)."""
    import os
    import numpy as np
    import torch
    from oracle import loss_oracle as lo
    from dynmm_b200.fusion import CrossEntropyLoss2d
    gold = np.load(os.path.join(golden_dir, "loss_ce2d.npz"))
    weight = gold["weight"]
    logits = [gold[f"logits{i}"] for i in range(4)]
    targets = [gold[f"targets{i}"] for i in range(4)]
    ref = [float(gold[f"loss{i}"]) for i in range(4)]
    got = lo.ce2d(logits, targets, weight)
    np.testing.assert_allclose(got, ref, rtol=2e-6)
    np.testing.assert_allclose(lo.ce2d_scale(gold["edge_logits"], gold["edge_targets"], weight), float(gold["edge_loss"]),
                               rtol=2e-6)
    loss_fn = CrossEntropyLoss2d(torch.device("cpu"), weight)
    xs = [torch.from_numpy(x).requires_grad_(True) for x in logits]
    losses = loss_fn(xs, [torch.from_numpy(t) for t in targets])
    np.testing.assert_allclose([l.item() for l in losses], ref, rtol=2e-6)
    sum(losses).backward()
    for i, x in enumerate(xs):
        np.testing.assert_allclose(x.grad.numpy(), gold[f"grad{i}"], rtol=1e-5, atol=1e-9)
    edge = loss_fn([torch.from_numpy(gold["edge_logits"])], [torch.from_numpy(gold["edge_targets"])])[0].item()
    np.testing.assert_allclose(edge, float(gold["edge_loss"]), rtol=2e-6)
