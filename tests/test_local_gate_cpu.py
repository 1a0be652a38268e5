"""Local-gate dynamic ESANet (dynmm_b200/fusion/local_gate.py) against vectors produced by the
reference's own ``SkipESANet`` (oracle/make_golden_local.py -> tests/golden/local_gate_*.npz): state_dict keys,
Gumbel / random-policy gate decisions (bit-exact: same generator, same order of draws), chained weights, logits in
every block_rule / mode combination, training-mode outputs; plus the build_model() mapping and the statistics API."""
import argparse
import os
import warnings

import numpy as np
import pytest
import torch

from oracle.make_golden_local import CASES, MODES, apply_mode, sample_inputs, seeded_state

warnings.filterwarnings("ignore")


def _model(name):
    from dynmm_b200.fusion import SkipESANet
    kw, seed, b = CASES[name]
    model = SkipESANet(pretrained_on_imagenet=False, **kw)
    model.load_state_dict(seeded_state(model.state_dict(), seed), strict=True)
    return model, sample_inputs(seed + 100, b, kw["height"], kw["width"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_local_gate_matches_reference_vectors(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    model, (rgb, depth) = _model(name)
    assert sorted(model.state_dict().keys()) == list(gold["keys"])          # strict load into the REFERENCE succeeded
    model.eval()
    with torch.no_grad():
        for tag, rule, attrs, test, fseed in MODES:
            apply_mode(model, rule, attrs)
            model.start_weight()
            torch.manual_seed(fseed)
            out = model(rgb, depth, test)
            model._flush_weights()
            for i in range(4):
                ref_w = gold[f"{tag}_weight{i}"]
                got_w = model.weight_list[i].numpy()
                hard = test or attrs.get("hard_gate", False) or attrs.get("random_policy", False)
                if hard:
                    np.testing.assert_array_equal(got_w, ref_w, err_msg=f"{tag} gate {i}")     # decisions bit-exact
                else:
                    np.testing.assert_allclose(got_w, ref_w, rtol=1e-5, atol=1e-6, err_msg=f"{tag} gate {i}")
            model.end_weight()
            ref = gold[f"{tag}_out"]
            np.testing.assert_allclose(out[:, :, ::4, ::4].numpy(), ref, rtol=2e-4, atol=2e-4 * np.abs(ref).max(),
                                       err_msg=tag)
            assert abs(out.double().abs().sum().item() - gold[f"{tag}_abssum"]) <= 2e-4 * gold[f"{tag}_abssum"]
    model.train()
    apply_mode(model, [2, 2, 2, 2], {})
    torch.manual_seed(21)
    with torch.no_grad():
        outs = model(rgb, depth)
    assert len(outs) == 4
    for i, o in enumerate(outs):
        assert list(o.shape) == list(gold[f"train_out{i}_shape"])
        assert abs(o.double().abs().sum().item() - gold[f"train_out{i}_abssum"]) <= 5e-4 * gold[f"train_out{i}_abssum"]
    ref = gold["train_out0"]
    np.testing.assert_allclose(outs[0][:, :, ::4, ::4].numpy(), ref, rtol=5e-4, atol=5e-4 * np.abs(ref).max())


def test_gate_needs_only_the_pooled_features():
    """mean(x * SE(x)) == mean_c(SE(gap)_c * gap_c): the module never builds cat(rgb, depth) or its scaled copy."""
    from dynmm_b200.fusion import SqueezeAndExcitationWeight
    torch.manual_seed(3)
    se = SqueezeAndExcitationWeight(32)
    x = torch.randn(4, 32, 9, 7)
    w = torch.nn.functional.adaptive_avg_pool2d(x, 1)
    ref = (x * se.fc(w).expand_as(x)).mean(dim=(1, 2, 3))                  # model_utils.py:66-70
    np.testing.assert_allclose(se(x).detach().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_gradients_reach_only_the_gates_after_freeze():
    model, (rgb, depth) = _model("local_gate_r18_basic_64x64")
    model.train()
    model.freeze()                                                         # model_skip_mod.py:215-218
    apply_mode(model, [2, 2, 2, 2], {})
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all("gate" in n for n in trainable)
    torch.manual_seed(0)
    outs = model(rgb, depth)
    outs[0].square().mean().backward()
    used = [n for n, p in model.named_parameters() if p.grad is not None and p.grad.abs().sum() > 0]
    assert used and all("gate_layer" in n and ".se.fc." in n for n in used)
    assert not any(".linear." in n for n in used)                           # the reference's unused head stays unused


def test_statistics_api_and_build_model():
    model, (rgb, depth) = _model("local_gate_r18_basic_64x64")
    model.eval()
    apply_mode(model, [1, 2, 2, 1], dict(hard_gate=True))
    model.start_weight()
    with torch.no_grad():
        for s in range(3):
            torch.manual_seed(s)
            model(rgb, depth, True)
    avg = model.end_weight()                                               # one mean per DYNAMIC site (:228-230)
    assert len(avg) == 2 and all(a.shape == (2,) and abs(a.sum().item() - 1) < 1e-6 for a in avg)
    assert all(w.numel() == 0 for w in model.weight_list) and model.save_weight_info is False

    from dynmm_b200.fusion import SkipESANet, SkipGateESANet, build_model
    args = argparse.Namespace(dynamic=True, global_gate=False, block_rule="1122", height=64, width=64,
                              encoder="resnet18", encoder_depth=None, encoder_block="BasicBlock", activation="relu",
                              encoder_decoder_fusion="add", context_module="ppm", nr_decoder_blocks=[1],
                              channels_decoder=128, decoder_channels_mode="constant",
                              fuse_depth_in_rgb_encoder="SE-add", upsampling="bilinear", temp=0.5,
                              pretrained_on_imagenet=False, last_ckpt="", pretrained_scenenet="", pretrained_dir="",
                              he_init=False, finetune=None)
    m, _ = build_model(args, n_classes=37)
    assert isinstance(m, SkipESANet) and m.block_rule == [1, 1, 2, 2] and m.gate_layer2.temp == 0.5
    args.global_gate = True
    m, _ = build_model(args, n_classes=37)
    assert isinstance(m, SkipGateESANet)
    args.dynamic = False
    with pytest.raises(NotImplementedError):
        build_model(args, n_classes=37)
