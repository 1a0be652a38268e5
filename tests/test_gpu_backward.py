"""Backward of the hot path on the GPU.

* weight gradient: the tcgen05 pixel-K GEMM against (i) an fp32 PyTorch reference fed the same bf16-rounded
  operands (`torch.nn.grad.conv2d_weight`) -- fp32 accumulation on both sides, tolerance 1e-3 of the gradient
  scale -- and (ii) at full NYUv2 sizes an independent CUDA-core kernel;
* data gradient (forward kernel, mirrored taps, zero-interleaved dy for stride 2) against
  `torch.nn.grad.conv2d_input`, bf16 output: 2 bf16 ulps + 2e-3;
* a whole training step with `train_precision="bf16"` (all encoder/decoder convolutions forward AND backward
  on our kernels) against the fp32 reference graph with frozen BatchNorm statistics: loss within 1e-2
  relative, every parameter gradient with cosine similarity > 0.95 (bf16 activations vs fp32); with batch
  statistics (an ill-conditioned comparison at random initialisation, see the test) loss and gradient norms.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


# name, n, h, w, c_in, c_out, (kh, kw), (sh, sw)
WGRAD_CASES = [
    ("3x1_c64", 2, 24, 40, 64, 64, (3, 1), (1, 1)),
    ("1x3_c128", 2, 24, 40, 128, 128, (1, 3), (1, 1)),
    ("3x1_s2_64to128", 2, 24, 40, 64, 128, (3, 1), (2, 1)),
    ("1x3_s2_c128", 2, 12, 40, 128, 128, (1, 3), (1, 2)),
    ("1x1_s2_64to128", 2, 24, 40, 64, 128, (1, 1), (2, 2)),
    ("3x3_256to128", 2, 15, 20, 256, 128, (3, 3), (1, 1)),
    ("3x3_128to40", 1, 30, 40, 128, 40, (3, 3), (1, 1)),
    ("1x1_128to40", 3, 15, 20, 128, 40, (1, 1), (1, 1)),
    ("3x1_c512_odd", 3, 15, 20, 512, 512, (3, 1), (1, 1)),
    ("1x3_c256_tiny", 1, 5, 7, 256, 256, (1, 3), (1, 1)),
    ("3x3_s2_64to128", 2, 17, 23, 64, 128, (3, 3), (2, 2)),
]


def _case_tensors(n, h, w, ci, co, k, s, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = (k[0] // 2, k[1] // 2)
    ho = (h + 2 * pad[0] - k[0]) // s[0] + 1
    wo = (w + 2 * pad[1] - k[1]) // s[1] + 1
    x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(n, ho, wo, co, device="cuda", generator=g).to(torch.bfloat16)
    return x, dy, pad


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad_matches_fp32_reference(case):
    from dynmm_b200 import ops
    _, n, h, w, ci, co, k, s = case
    x, dy, pad = _case_tensors(n, h, w, ci, co, k, s)
    got = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], stride=s, pad=pad)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (co, ci, k[0], k[1]),
                                      dy.float().permute(0, 3, 1, 2), stride=s, padding=pad)
    scale = ref.abs().max().item()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-3 * scale, (got - ref).abs().max().item() / scale
    # comparator kernel agrees too, and accumulate=True adds
    direct = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], stride=s, pad=pad, direct=True)
    assert (direct - ref).abs().max().item() < 1e-3 * scale
    acc = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], stride=s, pad=pad, out=got.clone(), accumulate=True)
    assert (acc - 2 * ref).abs().max().item() < 2e-3 * scale
    # deterministic: bit-identical on a second run
    again = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], stride=s, pad=pad)
    assert torch.equal(got, again)


@pytest.mark.parametrize("shape", [(8, 120, 160, 64, (1, 3)), (8, 60, 80, 128, (3, 1)), (8, 30, 40, 256, (1, 3)),
                                   (8, 15, 20, 512, (3, 1))])
def test_conv_wgrad_full_size_against_direct_kernel(shape):
    """Encoder layer shapes at 480x640, batch 8: tensor-core kernel vs the CUDA-core comparator."""
    from dynmm_b200 import ops
    n, h, w, c, k = shape
    x, dy, pad = _case_tensors(n, h, w, c, c, k, (1, 1), seed=3)
    got = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], pad=pad)
    ref = ops.conv_wgrad(x, dy, kh=k[0], kw=k[1], pad=pad, direct=True)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2e-3 * scale


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_dgrad_matches_fp32_reference(case):
    from dynmm_b200.fusion.train_ops import conv2d
    _, n, h, w, ci, co, k, s = case
    x, dy, pad = _case_tensors(n, h, w, ci, co, k, s, seed=1)
    wt = (torch.randn(co, ci, k[0], k[1], device="cuda") * (ci * k[0] * k[1]) ** -0.5).requires_grad_()
    bias = torch.randn(co, device="cuda").requires_grad_()
    xin = x.permute(0, 3, 1, 2).requires_grad_()
    y = conv2d(xin, wt, bias, s, pad)
    assert y.dtype == torch.bfloat16 and y.shape[1] == co
    y.backward(dy.permute(0, 3, 1, 2))
    wb = wt.detach().to(torch.bfloat16).float()
    y_ref = F.conv2d(x.float().permute(0, 3, 1, 2), wb, bias.detach(), s, pad)
    err = (y.float() - y_ref).abs()
    assert not (err > 2.0 ** -7 * y_ref.abs() + 2e-3).any()
    gx_ref = torch.nn.grad.conv2d_input(xin.shape, wb, dy.float().permute(0, 3, 1, 2), stride=s, padding=pad)
    err = (xin.grad.float() - gx_ref).abs()
    tol = 2.0 ** -7 * gx_ref.abs() + 2e-3 * max(1.0, gx_ref.abs().max().item())
    assert not (err > tol).any(), f"dgrad max err {err.max().item():.4g} of {gx_ref.abs().max().item():.3g}"
    gw_ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), wt.shape, dy.float().permute(0, 3, 1, 2),
                                         stride=s, padding=pad)
    assert (wt.grad - gw_ref).abs().max().item() < 1e-3 * gw_ref.abs().max().item()
    gb_ref = dy.float().sum((0, 1, 2))
    assert torch.allclose(bias.grad, gb_ref, rtol=1e-4, atol=1e-3)


def test_weight_pack_and_channel_sum():
    from dynmm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    for co, ci, kh, kw in [(64, 64, 3, 1), (40, 128, 3, 3), (128, 64, 1, 1), (256, 72, 1, 3)]:
        w = torch.randn(co, ci, kh, kw, device="cuda", generator=g)
        fwd, dgr = ops.pack_conv_weight_pair(w)
        assert torch.equal(fwd, ops.pack_conv_weight(w))
        assert torch.equal(dgr, ops.pack_conv_weight_dgrad(w))
    for shape in [(8, 120, 160, 64), (3, 15, 20, 40), (2, 7, 9, 512), (1, 1, 3, 8)]:
        x = torch.randn(*shape, device="cuda", generator=g).to(torch.bfloat16)
        got = ops.channel_sum(x)
        ref = x.float().sum((0, 1, 2))
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-2), (shape, (got - ref).abs().max().item())
        assert torch.equal(got, ops.channel_sum(x))                      # deterministic
        acc = ops.channel_sum(x, out=got.clone(), accumulate=True)
        assert torch.allclose(acc, 2 * ref, rtol=1e-4, atol=2e-2)
    wide = torch.randn(4, 6, 6, 128, device="cuda", generator=g).to(torch.bfloat16)
    assert torch.allclose(ops.channel_sum(wide, c=40), wide[..., :40].float().sum((0, 1, 2)), rtol=1e-4, atol=1e-2)


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _train_step(sd, hw, batch, precision, bn_batch_stats, seed=1):
    from oracle.make_golden import sample_inputs
    from dynmm_b200.fusion import SkipGateESANet
    hh, ww = hw
    rgb, depth = (t.cuda() for t in sample_inputs(seed, batch, hh, ww))
    target = torch.randint(0, 40, (batch, hh, ww), device="cuda", generator=torch.Generator("cuda").manual_seed(9))
    m = SkipGateESANet(height=hh, width=ww).cuda()
    m.load_state_dict(sd)
    m.train()
    if not bn_batch_stats:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eval()
    m.temp, m.hard_gate, m.train_precision = 1.0, True, precision
    (o, s8, s16, s32), lf = m(rgb, depth)
    loss = F.cross_entropy(o, target) + 0.1 * lf
    for so in (s8, s16, s32):
        loss = loss + F.cross_entropy(F.interpolate(so.float(), (hh, ww), mode="nearest"), target)
    loss.backward()
    return m, o.detach(), float(loss.detach())


def test_bf16_training_step_runs_on_kernels_and_tracks_fp32_graph():
    """Whole forward + backward on the kernels vs the fp32 reference graph.  BatchNorm layers use their running
    statistics here (fine-tuning with frozen statistics): with batch statistics this randomly initialised
    59-layer network amplifies ANY perturbation by ~7 % per layer (bf16 rounding of the first layer is 30 % by
    the decoder, for library bf16 convolutions just the same -- tools/debug_train_bf16.py), so gradients of two
    correct implementations decorrelate and nothing could be concluded.  Per-layer exactness of the three
    convolution passes is covered by the tests above; the batch-statistics step is checked below for sanity."""
    from oracle import fusion_oracle as fo
    from dynmm_b200 import ops
    cfg = fo.FusionConfig(height=64, width=96)
    sd = fo.make_state_dict(cfg, 0, gate_scale=40.0)
    calls = {"conv": 0, "wgrad": 0}
    real_conv, real_wgrad = ops.conv, ops.conv_wgrad

    def count_conv(*a, **k):
        calls["conv"] += 1
        return real_conv(*a, **k)

    def count_wgrad(*a, **k):
        calls["wgrad"] += 1
        return real_wgrad(*a, **k)
    ops.conv, ops.conv_wgrad = count_conv, count_wgrad
    try:
        m16, out16, loss16 = _train_step(sd, (64, 96), 4, "bf16", bn_batch_stats=False)
    finally:
        ops.conv, ops.conv_wgrad = real_conv, real_wgrad
    m32, out32, loss32 = _train_step(sd, (64, 96), 4, "fp32", bn_batch_stats=False)
    # every eligible convolution went through the kernels: forward + dgrad launches, one wgrad per layer
    assert calls["wgrad"] >= 150 and calls["conv"] >= 2 * calls["wgrad"] - 4, calls
    p16, p32 = dict(m16.named_parameters()), dict(m32.named_parameters())
    report = {"loss16": loss16, "loss32": loss32, "out_rel": ((out16 - out32).norm() / out32.norm()).item()}
    worst = ("", 1.0)
    n_checked = 0
    for nme, p in p32.items():
        if p.grad is None or p.grad.norm().item() < 1e-6:
            continue
        g16 = p16[nme].grad
        assert g16 is not None and torch.isfinite(g16).all(), nme
        if p.grad.numel() < 64:
            continue                       # few-element gradients: cosine is not informative
        c = _cos(g16, p.grad)
        n_checked += 1
        if c < worst[1]:
            worst = (nme, c)
    report["worst_cos"], report["checked"] = worst, n_checked
    print(report)
    assert abs(loss16 - loss32) < 1e-2 * abs(loss32), report
    assert report["out_rel"] < 2e-2, report
    assert n_checked > 300 and worst[1] > 0.95, report


def test_bf16_training_step_with_batch_statistics_is_sane():
    from oracle import fusion_oracle as fo
    cfg = fo.FusionConfig(height=160, width=224)
    sd = fo.make_state_dict(cfg, 0, gate_scale=40.0)
    m16, out16, loss16 = _train_step(sd, (160, 224), 4, "bf16", bn_batch_stats=True)
    m32, out32, loss32 = _train_step(sd, (160, 224), 4, "fp32", bn_batch_stats=True)
    assert abs(loss16 - loss32) < 2e-2 * abs(loss32), (loss16, loss32)
    p16, p32 = dict(m16.named_parameters()), dict(m32.named_parameters())
    for nme, p in p32.items():
        if p.grad is None:
            continue
        g16 = p16[nme].grad
        assert g16 is not None and torch.isfinite(g16).all(), nme
        # (a bias in front of a batch-statistics BatchNorm has an exactly-zero gradient: only rounding noise)
        # With batch statistics this network decorrelates the gradients of two correct implementations (see the test
        # above; cosine ~ 0 for most layers, tools/debug_gate_grad.py), so only magnitudes are compared, loosely.  The
        # gate's gradient is left out: with hard decisions it is ONE direction times a scalar that sums decorrelated
        # contributions of the four blend sites -- its norm ratio came out as 0.43, 0.33 and 4.1 for the same code on
        # three boxes, with either stem layout.
        if p.grad.numel() >= 64 and p.grad.norm().item() > 1e-3 and not nme.startswith("gate_layer."):
            ratio = (g16.norm() / p.grad.norm()).item()
            assert 0.33 < ratio < 3.0, (nme, ratio)
    # the layers next to the loss have not accumulated any amplification yet
    assert _cos(p16["decoder.conv_out.weight"].grad, p32["decoder.conv_out.weight"].grad) > 0.99
    for nme in ("encoder_rgb.layer1.0.bn1.running_mean", "encoder_depth.layer4.2.bn2.running_var"):
        b16, b32 = dict(m16.named_buffers())[nme], dict(m32.named_buffers())[nme]
        assert torch.allclose(b16, b32, rtol=0.1, atol=0.05), nme


def test_bf16_frozen_training_only_needs_data_gradients():
    """`freeze()` (model_skip_mod_globalgate.py:225-228): only gate parameters train, so the encoders
    contribute data gradients only -- no weight-gradient launch at all."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    from dynmm_b200 import ops
    from dynmm_b200.fusion import SkipGateESANet
    cfg = fo.FusionConfig(height=64, width=96)
    sd = fo.make_state_dict(cfg, 0, gate_scale=40.0)
    rgb, depth = sample_inputs(2, 2, 64, 96)
    m = SkipGateESANet(height=64, width=96).cuda()
    m.load_state_dict(sd)
    m.train()
    m.freeze()
    m.temp, m.hard_gate, m.train_precision = 1.0, False, "bf16"
    n_wgrad = [0]
    real = ops.conv_wgrad

    def counting(*a, **k):
        n_wgrad[0] += 1
        return real(*a, **k)
    ops.conv_wgrad = counting
    try:
        (o, *_), lf = m(rgb.cuda(), depth.cuda())
        (o.float().square().mean() + lf).backward()
    finally:
        ops.conv_wgrad = real
    assert n_wgrad[0] == 0
    g = m.gate_layer.fc.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    assert m.encoder_rgb.layer1[0].conv3x1_1.weight.grad is None
