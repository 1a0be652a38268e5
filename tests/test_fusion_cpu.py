"""Host-side logic of the fusion drop-in: state_dict compatibility with the reference
(key list stored in the golden files by a strict load into the REFERENCE model), the
differentiable PyTorch graph against reference-generated vectors, build_model()."""
import argparse
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle.make_golden import sample_inputs
from tests.test_oracle_golden import CASES

warnings.filterwarnings("ignore")


def _model(cfg):
    from dynmm_b200.fusion import SkipGateESANet
    return SkipGateESANet(height=cfg.height, width=cfg.width, encoder_rgb=cfg.encoder, encoder_depth=cfg.encoder,
                          encoder_block=cfg.encoder_block, channels_decoder=list(cfg.channels_decoder),
                          nr_decoder_blocks=list(cfg.nr_decoder_blocks),
                          fuse_depth_in_rgb_encoder=cfg.fuse_depth_in_rgb_encoder)


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_and_torch_path_match_reference(name, golden_dir):
    cfg, seed, b = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    model = _model(cfg)
    assert sorted(model.state_dict().keys()) == list(gold["keys"])
    sd = fo.make_state_dict(cfg, seed, float(gold["gate_scale"]))
    model.load_state_dict(sd, strict=True)
    model.eval()
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    with torch.no_grad():
        model.hard_gate = True
        out, w = model(rgb, depth, True, True)
    np.testing.assert_array_equal(w.numpy(), gold["hard_t1_weight"])
    ref = gold["hard_t1_out_sample"]
    np.testing.assert_allclose(out[:, :, ::4, ::4].numpy(), ref, rtol=2e-4, atol=2e-4 * np.abs(ref).max())
    model.train()
    model.hard_gate = False
    outs, loss = model(rgb, depth)
    assert len(outs) == 4 and abs(loss.item() - float(gold["train_loss"])) < 1e-5 * abs(float(gold["train_loss"])) + 1e-7
    for i, o in enumerate(outs):
        assert list(o.shape) == list(gold[f"train_out{i}_shape"])
        assert abs(o.double().abs().sum().item() - gold[f"train_out{i}_abssum"]) <= 5e-4 * gold[f"train_out{i}_abssum"]


def test_mode_attributes_and_freeze():
    cfg = fo.FusionConfig(height=64, width=64)
    model = _model(cfg)
    for attr in ("temp", "hard_gate", "baseline", "ini_stage", "save_weight_info", "weight_list", "flop",
                 "depth_enc_flop", "total_flop"):
        assert hasattr(model, attr)
    assert torch.allclose(model.total_flop, torch.tensor(fo.TOTAL_FLOP_R34))
    model.freeze()
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all("gate" in n for n in trainable)
    model.eval()
    model.start_weight()
    rgb, depth = sample_inputs(1, 2, 64, 64)
    with torch.no_grad():
        model.baseline = True
        model(rgb, depth, True)
    stats = model.end_weight(print_flop=True)
    assert stats[0].tolist() == [0, 0, 0, 0, 2]
    assert abs(stats[2] - fo.TOTAL_FLOP_R34[4]) < 1e-4


def test_build_model_signature():
    from dynmm_b200.fusion import build_model
    args = argparse.Namespace(dynamic=True, global_gate=True, block_rule="1111", height=64, width=64,
                              encoder="resnet34", encoder_depth=None, encoder_block="NonBottleneck1D", activation="relu",
                              encoder_decoder_fusion="add", context_module="ppm", nr_decoder_blocks=[3],
                              channels_decoder=128, decoder_channels_mode="constant", fuse_depth_in_rgb_encoder="add",
                              upsampling="learned-3x3-zeropad", temp=1.0, pretrained_on_imagenet=False, last_ckpt="",
                              pretrained_scenenet="", pretrained_dir="", he_init=True, finetune=None)
    model, device = build_model(args, n_classes=40)
    assert isinstance(device, torch.device) and model.decoder.conv_out.out_channels == 40
    args.global_gate = False                                   # build_model.py:76-95: the local-gate variant
    from dynmm_b200.fusion import SkipESANet
    assert isinstance(build_model(args, 40)[0], SkipESANet)
    args.dynamic = False                                       # static / one-modality ESANets are not provided
    with pytest.raises(NotImplementedError):
        build_model(args, 40)
    with pytest.raises(NotImplementedError):
        from dynmm_b200.fusion import SkipGateESANet
        SkipGateESANet(encoder_rgb="vgg16")


def test_he_init_follows_the_reference_rules():
    """build_model.py:152-178: output layers, SE convs followed by a Sigmoid and depthwise convs keep their init, no
    bias is touched, encoders and gate ARE re-initialised (no ImageNet weights), BatchNorm -> (1, 0)."""
    from dynmm_b200.fusion import SkipGateESANet
    from dynmm_b200.fusion.build import he_init
    torch.manual_seed(3)
    model = SkipGateESANet(height=64, width=64, num_classes=40, encoder_rgb="resnet18", encoder_depth="resnet18",
                           encoder_block="BasicBlock", fuse_depth_in_rgb_encoder="SE-add")
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.fill_(0.5)
                m.bias.fill_(0.25)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    he_init(model, 40, pretrained_on_imagenet=False)
    after = model.state_dict()
    same = lambda k: torch.equal(before[k], after[k])
    assert same("decoder.conv_out.weight") and same("decoder.decoder_module_1.side_output.weight")       # out == n_classes
    assert same("se_layer1.se_rgb.fc.2.weight") and not same("se_layer1.se_rgb.fc.0.weight")            # before Sigmoid
    assert same("decoder.upsample1.conv.weight") and same("decoder.decoder_module_1.upsample.conv.weight")   # depthwise
    conv_biases = [n + ".bias" for n, m in model.named_modules() if isinstance(m, torch.nn.Conv2d) and m.bias is not None]
    assert len(conv_biases) > 10 and all(same(k) for k in conv_biases), "conv biases must not be touched"
    assert not same("encoder_rgb.layer1.0.conv1.weight") and not same("gate_layer.conv.0.weight")
    assert not same("gate_layer.fc.weight")        # last module: the reference's unguarded index would raise here
    assert float(after["encoder_depth.bn1.weight"].min()) == 1.0 and float(after["encoder_depth.bn1.bias"].abs().max()) == 0.0
    before2 = {k: v.clone() for k, v in after.items()}
    he_init(model, 40, pretrained_on_imagenet=True)      # ImageNet encoders are skipped
    assert torch.equal(before2["encoder_rgb.layer1.0.conv1.weight"], model.state_dict()["encoder_rgb.layer1.0.conv1.weight"])


def test_conv_program_recorder_phases():
    """ConvProgram (host side of dynmm_conv_program_*): phases are numbered without gaps, at most 4 jobs each."""
    import pytest
    from dynmm_b200 import ops, _lib
    prog = ops.ConvProgram()
    prog.next_phase()                      # an empty phase is not numbered
    assert prog.phase == 0
    for _ in range(3):
        prog._record(_lib.ConvParams(), (None,) * 12)
    assert prog.jobs_in_phase() == 3
    prog.next_phase()
    prog.next_phase()                      # second call: the new phase is still empty
    assert prog.phase == 1
    for _ in range(4):
        prog._record(_lib.ConvParams(), (None,) * 12)
    with pytest.raises(_lib.DynmmError):
        prog._record(_lib.ConvParams(), (None,) * 12)
    assert prog.phases == [0, 0, 0, 1, 1, 1, 1]


def test_chain_layer_lists_follow_the_block_structure():
    """ops.nbt1d_chain_layers (host logic of dynmm_conv_chain_*): taps alternate 3x1 / 1x3, the fourth convolution of
    a block adds the block input (chain input for the first block, the stored block output afterwards) and stores its
    output; with drop_last the gated last convolution is left out and the layer before it is stored separately."""
    from dynmm_b200 import ops
    blocks = [[(f"w{b}{i}", f"s{b}{i}", True) for i in range(4)] for b in range(3)]
    full = ops.nbt1d_chain_layers(blocks)
    assert len(full) == 12
    assert [l[2] for l in full] == [True, False] * 6                       # taps_h
    assert [l[4] for l in full] == [0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 2]    # residual source
    assert [l[5] for l in full] == [0, 0, 0, 1] * 3                        # stored to `out`
    cut = ops.nbt1d_chain_layers(blocks, drop_last=True)
    assert len(cut) == 11 and cut[-1][0] == "w22" and cut[-1][5] == 2 and cut[-1][4] == 0
    assert [l[5] for l in cut[:-1]] == [0, 0, 0, 1, 0, 0, 0, 1, 0, 0]
    one = ops.nbt1d_chain_layers(blocks[:1], drop_last=True)               # a two-block stage: nothing stored to `out`
    assert [l[5] for l in one] == [0, 0, 2]


def test_stem_bn_host_vector_layout():
    """The host copy of the stem's BN vectors that travels as a kernel parameter (dynmm_stem_s2d_fwd's ``bn_host``):
    [scale_rgb | shift_rgb | scale_d | shift_d], 256 contiguous fp32 values on the host."""
    from dynmm_b200 import ops
    parts = [torch.arange(64, dtype=torch.float32) + 100 * i for i in range(4)]
    v = ops.stem_s2d_bn_host(parts[0], parts[1].double(), parts[2], parts[3].reshape(64, 1))
    assert v.dtype == torch.float32 and not v.is_cuda and v.is_contiguous() and v.shape == (256,)
    assert torch.equal(v, torch.cat([p.reshape(64).float() for p in parts]))


def test_training_stem_switches_to_channels_last_only_for_the_bf16_graph(monkeypatch):
    """`stem_channels_last` (bf16 training graph on CUDA) is a layout change only: same values as forward_first_conv."""
    from dynmm_b200.fusion.modules import stem_channels_last
    m = _model(fo.FusionConfig(height=64, width=64))
    m.train()
    x = torch.randn(2, 3, 64, 64)
    torch.manual_seed(0)
    ref = m.encoder_rgb.forward_first_conv(x)
    got = stem_channels_last(m.encoder_rgb, x)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    monkeypatch.setenv("DYNMM_TRAIN_STEM", "nchw")
    assert stem_channels_last(m.encoder_rgb, x).is_contiguous()
