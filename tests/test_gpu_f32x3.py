"""fp32-grade engine mode ("f32x3"): activations and weights as bf16 hi + lo halves, three tensor-core products per
MAC (DYNMM_CONV_SPLIT), element-wise kernels in fp32.  north_star's fp32 bar: logits within 1e-3 relative of the
reference's fp32 path (tests/golden: produced by the reference's SkipGateESANet), hard decisions bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

F32_TOL = 1e-3          # north_star: "logits ... within 1e-3 relative for fp32"


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def _join(t):
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


def test_split_roundtrip_keeps_16_bits():
    from dynmm_b200 import ops
    x = torch.randn(3, 5, 7, 64, device="cuda") * 3
    s = ops.split_from_f32(x)
    assert s.shape == (3, 5, 7, 128) and s.dtype == torch.bfloat16
    assert torch.equal(s[..., :64], x.to(torch.bfloat16))
    assert (_join(s) - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()


@pytest.mark.parametrize("c_in,c_out,k,stride,h,w,res,gated", [
    (128, 128, (3, 1), (1, 1), 30, 40, True, False),      # halo mode, residual
    (64, 128, (3, 3), (2, 2), 32, 48, False, False),      # strided, per-tap loads
    (256, 256, (1, 3), (1, 1), 15, 20, True, True),       # gated depth add
    (128, 40, (3, 3), (1, 1), 24, 32, False, False),      # conv_out: 40 classes
    (512, 128, (1, 1), (1, 1), 5, 5, False, False),       # pyramid-pooling branch
])
def test_split_conv_matches_fp32(c_in, c_out, k, stride, h, w, res, gated):
    from dynmm_b200 import ops
    dev = "cuda"
    torch.manual_seed(c_in + c_out)
    n = 3
    kh, kw = k
    x = torch.randn(n, c_in, h, w, device=dev)
    wt = torch.randn(c_out, c_in, kh, kw, device=dev) / (c_in * kh * kw) ** 0.5
    bias = torch.randn(c_out, device=dev) * 0.1
    pad = (kh // 2, kw // 2)
    ref = F.conv2d(x, wt, bias, stride=stride, padding=pad)
    packed, shift = ops.fold_pack_conv(wt, bias, None, split=True)
    xs = ops.split_from_f32(x.permute(0, 2, 3, 1).contiguous())
    kw_ = {}
    if res:
        r = torch.randn_like(ref)
        ref = ref + r
        kw_["residual"] = ops.split_from_f32(r.permute(0, 2, 3, 1).contiguous())
    ref = torch.relu(ref)
    if gated:
        g = torch.tensor([0.0, 1.0, 0.37], device=dev)
        dpt = torch.randn_like(ref)
        ref = ref + g.view(-1, 1, 1, 1) * dpt
        kw_.update(gated=ops.split_from_f32(dpt.permute(0, 2, 3, 1).contiguous()), gate=g)
    out = ops.conv(xs, packed, c_out=c_out, kh=kh, kw=kw, stride=stride, pad=pad, shift=shift, relu=True, split=True, **kw_)
    got = _join(out).permute(0, 3, 1, 2)
    err = _rel_l2(got, ref)
    assert err < 2e-5, f"relative L2 {err:.2e}"
    assert (got - ref).abs().max().item() < 1e-3 * ref.abs().max().item()


def test_split_elementwise_kernels_match_fp32():
    from dynmm_b200 import ops
    dev = "cuda"
    torch.manual_seed(5)
    n, h, w, c = 2, 9, 12, 40
    x = torch.randn(n, c, h, w, device=dev)
    wt = torch.randn(c, 1, 3, 3, device=dev) * 0.3
    b = torch.randn(c, device=dev) * 0.1
    skip = torch.randn(n, c, 2 * h, 2 * w, device=dev)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wt, b, padding=1, groups=c)
    taps = wt.reshape(c, 9).t().contiguous()
    xs = ops.split_from_f32(x.permute(0, 2, 3, 1).contiguous())
    ss = ops.split_from_f32(skip.permute(0, 2, 3, 1).contiguous())
    got = _join(ops.upsample2x_dw3x3(xs, taps, b, ss, split=True)).permute(0, 3, 1, 2)
    assert _rel_l2(got, ref + skip) < 1e-5
    labels = torch.empty(n, 2 * h, 2 * w, dtype=torch.uint8, device=dev)
    out = ops.upsample2x_dw3x3(xs, taps, b, to_nchw_f32=True, labels=labels, split=True)
    assert _rel_l2(out, ref) < 1e-5
    assert torch.equal(labels.long(), out.argmax(1))
    # pyramid pooling helpers on a concat buffer [hi(96) | lo(96)]: the first 64 channels hold the features
    cat = torch.zeros(n, h, w, 192, dtype=torch.bfloat16, device=dev)
    feat = torch.randn(n, h, w, 64, device=dev)
    fs = ops.split_from_f32(feat)
    cat[..., :64], cat[..., 96:160] = fs[..., :64], fs[..., 64:]
    for bins in (1, 5):
        pooled = _join(ops.adaptive_avgpool(cat, bins, c=64, split=True)).permute(0, 3, 1, 2)
        ref_p = F.adaptive_avg_pool2d(_join(fs).permute(0, 3, 1, 2), bins)
        assert _rel_l2(pooled, ref_p) < 1e-5
    y = torch.randn(n, 5, 5, 32, device=dev)
    ys = ops.split_from_f32(y)
    ops.nearest_resize_into(ys, cat, 64, split=True)
    ref_y = F.interpolate(_join(ys).permute(0, 3, 1, 2), size=(h, w), mode="nearest").permute(0, 2, 3, 1)
    got_y = cat[..., 64:96].float() + cat[..., 160:192].float()
    assert torch.equal(got_y, ref_y)


def _build(cfg, seed, gate_scale=40.0):
    from dynmm_b200.fusion import SkipGateESANet
    from oracle import fusion_oracle as fo
    sd = fo.make_state_dict(cfg, seed, gate_scale)
    model = SkipGateESANet(height=cfg.height, width=cfg.width, encoder_rgb=cfg.encoder, encoder_depth=cfg.encoder,
                           encoder_block=cfg.encoder_block, channels_decoder=list(cfg.channels_decoder),
                           nr_decoder_blocks=list(cfg.nr_decoder_blocks),
                           fuse_depth_in_rgb_encoder=cfg.fuse_depth_in_rgb_encoder)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    model.engine_precision = "f32x3"
    return model, sd


def test_f32x3_engine_matches_reference_golden_vectors(golden_dir):
    """The reference's own fp32 outputs (tests/golden) to 1e-3 relative, hard decisions bit-exact."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=64, width=96)
    gold = np.load(os.path.join(golden_dir, "fusion_r34_nbt1d_add_64x96.npz"))
    model, sd = _build(cfg, 0, float(gold["gate_scale"]))
    rgb, depth = sample_inputs(1, 4, 64, 96)
    rgb, depth = rgb.cuda(), depth.cuda()
    errs = {}
    with torch.no_grad():
        for tag, temp, hard in (("soft_t1", 1.0, False), ("hard_t1", 1.0, True), ("soft_t01", 0.1, False)):
            model.temp, model.hard_gate = temp, hard
            out, w = model(rgb, depth, True, True)
            assert model.engine().split
            if hard:
                np.testing.assert_array_equal(w.cpu().numpy(), gold[tag + "_weight"])
            ref = torch.from_numpy(gold[tag + "_out_sample"])
            errs[tag] = _rel_l2(out[:, :, ::4, ::4].cpu(), ref)
        eng = model.engine()
        for k in range(5):
            wk = torch.eye(5)[torch.full((4,), k)].cuda()
            out, _ = eng.forward(rgb, depth, weight=wk)
            errs[f"branch{k}"] = _rel_l2(out[:, :, ::4, ::4].cpu(), torch.from_numpy(gold[f"branch{k}_out_sample"]))
    assert max(errs.values()) <= F32_TOL, errs
    print("f32x3 vs reference fp32:", {k: f"{v:.1e}" for k, v in errs.items()})


def test_split_se_kernels_match_fp32():
    """Squeeze sums and the gated SE blend on [hi | lo] tensors (dynmm_gap_partial_split, dynmm_se_gated_fuse_split)."""
    from dynmm_b200 import ops
    torch.manual_seed(3)
    dev = "cuda"
    n, h, w, c = 4, 9, 11, 64
    r32, d32 = torch.randn(n, h, w, c, device=dev), torch.randn(3, h, w, c, device=dev)
    rs, ds = ops.split_from_f32(r32), ops.split_from_f32(d32)
    part = ops.gap_partial(rs, split=True)
    assert part.shape == (n, 64, c)
    ref_sum = _join(rs).double().sum((1, 2))
    assert ((part.double().sum(1) - ref_sum).abs().max() / ref_sum.abs().max()).item() < 1e-6
    count = torch.tensor([2], dtype=torch.int32, device=dev)
    part_d = ops.gap_partial(ds, count=count, split=True)
    assert torch.allclose(part_d[:2].double().sum(1), _join(ds)[:2].double().sum((1, 2)), rtol=1e-5, atol=1e-4)
    sig_r, sig_d = torch.rand(n, c, device=dev), torch.rand(3, c, device=dev)
    gate = torch.tensor([0.0, 1.0, 0.25, 1.0], device=dev)
    slot = torch.tensor([2, 0, 1, 2], dtype=torch.int32, device=dev)
    wide = torch.zeros(n, h, w, 2 * (c + 32), dtype=torch.bfloat16, device=dev)      # e.g. the pyramid-pooling concat buffer
    for out in (None, wide):
        got = ops.se_gated_fuse(rs, ds, sig_r, sig_d, gate, slot, out=out, split=True)
        half = got.shape[-1] // 2
        val = got[..., :c].float() + got[..., half:half + c].float()
        g = gate.view(n, 1, 1, 1)
        dd = _join(ds)[slot.long()]
        ref = _join(rs) * (1 - g + g * sig_r.view(n, 1, 1, c)) + g * sig_d[slot.long()].view(n, 1, 1, c) * dd
        assert _rel_l2(val, ref) < 1e-5
        assert torch.equal(val[0], _join(rs)[0])          # a gated-off sample keeps its RGB features bit for bit


@pytest.mark.parametrize("name", ["fusion_r34_nbt1d_seadd_64x64", "fusion_r50_seadd_decr_64x64"])
def test_f32x3_engine_se_add_matches_reference_golden_vectors(name, golden_dir):
    """'SE-add' fusion (the reference's CLI default, args.py:159) in the fp32-grade mode: the reference's own fp32
    outputs to 1e-3 relative, hard decisions bit-exact."""
    from tests.test_oracle_golden import CASES
    from oracle.make_golden import sample_inputs
    cfg, seed, b = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    model, sd = _build(cfg, seed, float(gold["gate_scale"]))
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    rgb, depth = rgb.cuda(), depth.cuda()
    errs = {}
    with torch.no_grad():
        for tag, temp, hard in (("hard_t1", 1.0, True), ("soft_t1", 1.0, False)):
            model.temp, model.hard_gate = temp, hard
            out, w = model(rgb, depth, True, True)
            assert model.engine().split and model.engine().se is not None
            if hard:
                np.testing.assert_array_equal(w.cpu().numpy(), gold[tag + "_weight"])
            errs[tag] = _rel_l2(out[:, :, ::4, ::4].cpu(), torch.from_numpy(gold[tag + "_out_sample"]))
        eng = model.engine()
        for k in (0, 2, 4):
            wk = torch.eye(5)[torch.full((b,), k)].cuda()
            out, _ = eng.forward(rgb, depth, weight=wk)
            errs[f"branch{k}"] = _rel_l2(out[:, :, ::4, ::4].cpu(), torch.from_numpy(gold[f"branch{k}_out_sample"]))
    assert max(errs.values()) <= F32_TOL, errs
    print(name, "f32x3 vs reference fp32:", {k: f"{v:.1e}" for k, v in errs.items()})


def test_full_size_batch_8_both_precisions_match_oracle():
    """480x640, the whole batch of 8 (the bench workload's shape) against the fp32 CPU oracle: f32x3 logits 1e-3
    relative and (almost) every arg-max label; the bf16 engine on the same images within its stated 2e-2."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig()
    model, sd = _build(cfg, 0)
    model.hard_gate = True
    rgb, depth = sample_inputs(11, 8, 480, 640)
    with torch.no_grad():
        ref = fo.forward(sd, cfg, rgb, depth, hard_gate=True)
        out, w = model(rgb.cuda(), depth.cuda(), True, True)
        assert torch.equal(w.cpu(), ref["weight"])
        err = _rel_l2(out.cpu(), ref["out"])
        assert err <= F32_TOL, f"f32x3: relative L2 {err:.2e}"
        agree = (out.cpu().argmax(1) == ref["out"].argmax(1)).float().mean().item()
        assert agree >= 0.9999, agree
        per_sample = [_rel_l2(out[i].cpu(), ref["out"][i]) for i in range(8)]
        assert max(per_sample) <= F32_TOL, per_sample
        model.engine_precision = "bf16"
        out16, w16 = model(rgb.cuda(), depth.cuda(), True, True)
        assert not model.engine().split
        assert torch.equal(w16.cpu(), ref["weight"])
        err16 = _rel_l2(out16.cpu(), ref["out"])
        assert err16 <= 2e-2, f"bf16: relative L2 {err16:.2e}"
    print(f"480x640 x 8 vs oracle: f32x3 {err:.1e} (arg-max agreement {agree:.5f}), bf16 {err16:.1e}")


def test_ppm_1_2_4_8_context_module_on_the_engine():
    """context_module='ppm-1-2-4-8' (context_modules.py:28-38: four pooling branches) runs on the engine in both
    precisions; the comparison is the module's own fp32 PyTorch graph on the same weights (the oracle restates 'ppm')."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(3)
    model = SkipGateESANet(height=96, width=128, num_classes=40, context_module="ppm-1-2-4-8").cuda().eval()
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
        model.gate_layer.fc.weight.mul_(40.0)
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(5, 4, 96, 128))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)          # eval mode: (logits, gate weight)
        for precision, tol in (("f32x3", F32_TOL), ("bf16", 2e-2)):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            assert getattr(model, "_engine_unsupported", None) is None and model.engine().ppm_bins == (1, 2, 4, 8)
            assert torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{precision}: relative L2 {err:.2e}"


@pytest.mark.parametrize("mode", ["bilinear", "nearest", "learned-3x3"])
def test_other_upsampling_modes_on_the_engine(mode):
    """Upsample modes of model.py:360-410 other than the default 'learned-3x3-zeropad': 'learned-3x3' pads the
    up-sampled map by replication, 'bilinear' is that form with the fixed [1 2 1]^T [1 2 1] / 16 stencil (and bilinear
    interpolation of the pyramid-pooling branches), 'nearest' the identity stencil -- all through the same kernels, in
    both precisions, against the module's own fp32 PyTorch graph."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(7)
    model = SkipGateESANet(height=96, width=128, num_classes=40, upsampling=mode).cuda().eval()
    g = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            if mode == "learned-3x3" and name.endswith("upsample.conv") or name.endswith(("upsample1.conv", "upsample2.conv")):
                if hasattr(m, "weight") and m.weight is not None:
                    m.weight.add_(torch.randn(m.weight.shape, generator=g).to(m.weight.device) * 0.05)
                    m.bias.add_(torch.randn(m.bias.shape, generator=g).to(m.bias.device) * 0.05)
        model.gate_layer.fc.weight.mul_(40.0)
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(9, 3, 96, 128))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)
        for precision, tol in (("f32x3", F32_TOL), ("bf16", 2e-2)):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            assert getattr(model, "_engine_unsupported", None) is None, model._engine_unsupported
            assert torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{mode}/{precision}: relative L2 {err:.2e}"


def test_37_classes_run_on_the_engine():
    """num_classes % 8 != 0 (SUN RGB-D: 37): the engine carries 40 NHWC channels (zero weight rows, stencils and biases)
    and the final kernel emits -- and arg-maxes over -- the 37 real classes only."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(11)
    model = SkipGateESANet(height=64, width=96, num_classes=37).cuda().eval()
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
        # push the logits down so that some pixels have only negative ones: a padded class (exactly 0) must not win there
        model.decoder.upsample2.conv.bias.sub_(8.0)
        model.gate_layer.fc.weight.mul_(40.0)
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(2, 3, 64, 96))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)
        for precision, tol in (("f32x3", F32_TOL), ("bf16", 2e-2)):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            assert getattr(model, "_engine_unsupported", None) is None
            assert out.shape == (3, 37, 64, 96) and torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{precision}: relative L2 {err:.2e}"
            labels = model.predict_labels(rgb, depth)
            assert int(labels.max()) < 37
            assert (labels.long() == out.argmax(1)).float().mean().item() == 1.0
    assert (ref.max(1).values < 0).any(), "the scenario needs pixels whose real logits are all negative"


def test_encoder_decoder_fusion_none_on_the_engine():
    """encoder_decoder_fusion='None' (model.py:353-355: the decoder modules do not add the encoder skip tensors)."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(13)
    model = SkipGateESANet(height=64, width=96, num_classes=40, encoder_decoder_fusion="None").cuda().eval()
    g = torch.Generator().manual_seed(14)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(3, 3, 64, 96))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)
        for precision, tol in (("f32x3", F32_TOL), ("bf16", 2e-2)):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            assert getattr(model, "_engine_unsupported", None) is None and not model.engine().dec_fusion
            assert torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{precision}: relative L2 {err:.2e}"


def test_different_encoders_on_the_engine():
    """encoder_rgb != encoder_depth (ResNet-34 for RGB, ResNet-18 for depth: equal stage widths, different depths):
    the encoders run as independent launch sequences on two streams (no merged launches, no chain)."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(15)
    model = SkipGateESANet(height=64, width=96, num_classes=40, encoder_rgb="resnet34", encoder_depth="resnet18").cuda().eval()
    g = torch.Generator().manual_seed(16)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
        model.gate_layer.fc.weight.mul_(40.0)
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(4, 4, 64, 96))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)
        for precision, tol in (("f32x3", F32_TOL), ("bf16", 2e-2)):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            eng = model.engine()
            assert getattr(model, "_engine_unsupported", None) is None and not eng.same_encoders and not eng.use_merge
            assert torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{precision}: relative L2 {err:.2e}"


@pytest.mark.parametrize("activation", ["swish", "hswish"])
def test_swish_and_hswish_models_on_the_engine(activation):
    """activation='swish' / 'hswish' (model_utils.py:100-115): activation code in the convolution epilogue, per-layer
    launches, library stem; the bf16 engine against the module's own fp32 PyTorch graph (f32x3 fuses ReLU only: such a
    model falls back to the PyTorch graph in that mode)."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle.make_golden import sample_inputs
    torch.manual_seed(17)
    model = SkipGateESANet(height=64, width=96, num_classes=40, activation=activation).cuda().eval()
    g = torch.Generator().manual_seed(18)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
        model.gate_layer.fc.weight.mul_(40.0)
    model.hard_gate = True
    rgb, depth = (t.cuda() for t in sample_inputs(6, 3, 64, 96))
    with torch.no_grad():
        ref, w_ref = model._forward_torch(rgb, depth)
        for precision, tol in (("bf16", 2e-2),):
            model.engine_precision = precision
            out, w = model(rgb, depth, True, True)
            assert getattr(model, "_engine_unsupported", None) is None, model._engine_unsupported
            assert model.engine().act == {"swish": 2, "hswish": 3}[activation]
            assert torch.equal(w, w_ref)
            err = _rel_l2(out, ref)
            assert err <= tol, f"{activation}/{precision}: relative L2 {err:.2e}"
