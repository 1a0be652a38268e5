"""Kernel-level parity (GPU): every C-ABI op against plain fp32 PyTorch on the
same inputs.  bf16 tensor-core kernels are compared with an fp32 reference fed
the SAME bf16-rounded operands, so the only differences are accumulation order
and the final bf16 rounding: tolerance = 2 bf16 ulps (2^-7 relative) + 2e-3
absolute.  fp32 kernels: 1e-5 relative."""
import numpy as np
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


def _bf16_close(got, ref, what=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} off, max err {err.max().item():.4g} " \
                          f"(ref magnitude {ref.abs().max().item():.3g})"


def _ref_conv(x, w, stride, pad, scale, shift, residual, relu, gated, gate, slot):
    """fp32 reference on NHWC bf16 operands (returned NHWC fp32)."""
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), None, stride, pad)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = F.relu(y)
    if gated is not None:
        idx = slot.long() if slot is not None else torch.arange(y.shape[0], device=y.device)
        y = y + gate.view(-1, 1, 1, 1) * gated.float()[idx]
    return y


CONV_CASES = [
    # name, n, h, w, cin, cout, kh, kw, stride, pad
    ("1x3_c64", 2, 12, 40, 64, 64, 1, 3, (1, 1), (0, 1)),
    ("3x1_c128", 3, 15, 20, 128, 128, 3, 1, (1, 1), (1, 0)),
    ("3x1_s2_64to128", 2, 24, 32, 64, 128, 3, 1, (2, 1), (1, 0)),
    ("1x3_s2_128", 2, 12, 32, 128, 128, 1, 3, (1, 2), (0, 1)),
    ("1x1_s2_ds", 2, 24, 32, 64, 128, 1, 1, (2, 2), (0, 0)),
    ("3x3_c128to40", 2, 9, 14, 128, 40, 3, 3, (1, 1), (1, 1)),
    ("3x3_s2_basic", 2, 16, 24, 64, 128, 3, 3, (2, 2), (1, 1)),
    ("1x1_c1024", 2, 15, 20, 1024, 128, 1, 1, (1, 1), (0, 0)),
    ("1x3_c512_small", 5, 15, 20, 512, 512, 1, 3, (1, 1), (0, 1)),
    ("3x1_c256", 2, 30, 40, 256, 256, 3, 1, (1, 1), (1, 0)),
    ("odd_hw", 1, 7, 11, 64, 64, 3, 3, (1, 1), (1, 1)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("full_epilogue", [False, True])
def test_conv_igemm_matches_fp32_reference(case, full_epilogue):
    from dynmm_b200 import ops
    name, n, h, w, cin, cout, kh, kw, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 1000)
    dev = "cuda"
    x = torch.randn(n, h, w, cin, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(cout, cin, kh, kw, device=dev, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5
    ho = (h + 2 * pad[0] - kh) // stride[0] + 1
    wo = (w + 2 * pad[1] - kw) // stride[1] + 1
    kwargs = dict(scale=None, shift=None, residual=None, relu=False, gated=None, gate=None, slot=None)
    if full_epilogue:
        kwargs["scale"] = 0.5 + torch.rand(cout, device=dev, generator=g)
        kwargs["shift"] = torch.randn(cout, device=dev, generator=g) * 0.1
        kwargs["residual"] = torch.randn(n, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
        kwargs["relu"] = True
        kwargs["gated"] = torch.randn(n, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
        gate = torch.rand(n, device=dev, generator=g)
        gate[0] = 0.0
        kwargs["gate"] = gate
        kwargs["slot"] = torch.randperm(n, device=dev, generator=g).to(torch.int32)
    ref = _ref_conv(x, wt, stride, pad, **kwargs)
    packed = ops.pack_conv_weight(wt)
    common = dict(c_out=cout, kh=kh, kw=kw, stride=stride, pad=pad, scale=kwargs["scale"], shift=kwargs["shift"],
                  residual=kwargs["residual"], relu=kwargs["relu"], gated=kwargs["gated"], gate=kwargs["gate"],
                  gated_slot=kwargs["slot"])
    got = ops.conv(x, packed, **common)
    torch.cuda.synchronize()
    _bf16_close(got, ref, name)
    # the CUDA-core comparator obeys the same contract
    got_d = ops.conv(x, packed, direct=True, **common)
    _bf16_close(got_d, ref, name + "/direct")
    for tile_n in (16, 64, 128, 256):
        if tile_n <= (cout + 15) // 16 * 16:
            got_t = ops.conv(x, packed, tile_n=tile_n, **common)
            _bf16_close(got_t, ref, f"{name}/tile_n={tile_n}")


def test_conv_igemm_sample_indirection_skips_work():
    """count / in_map: only the first `count` output slots are produced, reading the mapped
    input samples; slots beyond count are left untouched (sentinel survives)."""
    from dynmm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    n, h, w, c = 6, 15, 20, 64
    x = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(c, c, 3, 1, device=dev, generator=g) * 0.1
    packed = ops.pack_conv_weight(wt)
    perm = torch.tensor([4, 2, 5, 0, 1, 3], dtype=torch.int32, device=dev)
    for cnt in (0, 1, 3, 6):
        count = torch.tensor([cnt], dtype=torch.int32, device=dev)
        out = torch.full((n, h, w, c), 7.0, dtype=torch.bfloat16, device=dev)
        ops.conv(x, packed, c_out=c, kh=3, kw=1, pad=(1, 0), in_map=perm, count=count, out=out, relu=True)
        ref = _ref_conv(x[perm.long()], wt, (1, 1), (1, 0), None, None, None, True, None, None, None)
        torch.cuda.synchronize()
        if cnt:
            _bf16_close(out[:cnt], ref[:cnt], f"count={cnt}")
        assert (out[cnt:].float() == 7.0).all()
    # multi-sample boxes (small maps) with a count that is not a multiple of the box
    x = torch.randn(8, 6, 10, 64, device=dev, generator=g).to(torch.bfloat16)
    count = torch.tensor([5], dtype=torch.int32, device=dev)
    out = torch.full((8, 6, 10, 64), 7.0, dtype=torch.bfloat16, device=dev)
    ops.conv(x, packed, c_out=64, kh=3, kw=1, pad=(1, 0), count=count, out=out)
    ref = _ref_conv(x, wt, (1, 1), (1, 0), None, None, None, False, None, None, None)
    _bf16_close(out[:5], ref[:5], "box_n>1")
    assert (out[5:].float() == 7.0).all()


def test_conv_igemm_channel_slices():
    """in_ld / out_ld pitches: read a channel prefix, write into a slice of a wider buffer."""
    from dynmm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(2, 10, 12, 192, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(64, 128, 1, 1, device=dev, generator=g) * 0.1
    out = torch.zeros(2, 10, 12, 256, dtype=torch.bfloat16, device=dev)
    ops.conv(x, ops.pack_conv_weight(wt), c_out=64, kh=1, kw=1, c_in=128, out=out, out_c_off=64)
    ref = _ref_conv(x[..., :128], wt, (1, 1), (0, 0), None, None, None, False, None, None, None)
    _bf16_close(out[..., 64:128], ref, "slice")
    assert (out[..., :64] == 0).all() and (out[..., 128:] == 0).all()


@pytest.mark.parametrize("shape", [(8, 120, 160, 64, (1, 3)), (8, 60, 80, 128, (3, 1)), (8, 30, 40, 256, (1, 3)),
                                   (8, 15, 20, 512, (3, 1))])
def test_conv_igemm_full_size_against_direct_kernel(shape):
    """BASELINE sizes (B=8 stage shapes): tensor-core kernel vs the independent CUDA-core kernel."""
    from dynmm_b200 import ops
    n, h, w, c, (kh, kw) = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(c)
    x = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(c, c, kh, kw, device=dev, generator=g) * (2.0 / (3 * c)) ** 0.5
    res = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    scale = 0.5 + torch.rand(c, device=dev, generator=g)
    shift = torch.randn(c, device=dev, generator=g) * 0.1
    packed = ops.pack_conv_weight(wt)
    kw_ = dict(c_out=c, kh=kh, kw=kw, pad=(kh // 2, kw // 2), scale=scale, shift=shift, residual=res, relu=True)
    a = ops.conv(x, packed, **kw_)
    b = ops.conv(x, packed, direct=True, **kw_)
    torch.cuda.synchronize()
    _bf16_close(a, b, "full-size")
    # linearity (size-independent property): conv(2x) - 2*conv(x) == 0 without the affine epilogue
    y1 = ops.conv(x, packed, c_out=c, kh=kh, kw=kw, pad=(kh // 2, kw // 2)).float()
    y2 = ops.conv((x.float() * 2).to(torch.bfloat16), packed, c_out=c, kh=kh, kw=kw, pad=(kh // 2, kw // 2)).float()
    assert torch.equal(y2, 2 * y1)      # exact: scaling by 2 commutes with every rounding step


def _stem_weights(sd, enc):
    from dynmm_b200 import ops
    w = sd[f"{enc}.conv1.weight"].permute(2, 3, 1, 0).contiguous().cuda()     # [7][7][cin][64]
    s, b = ops.fold_bn(sd[f"{enc}.bn1.weight"], sd[f"{enc}.bn1.bias"], sd[f"{enc}.bn1.running_mean"],
                       sd[f"{enc}.bn1.running_var"], 1e-5)
    return w, s.cuda(), b.cuda()


@pytest.mark.parametrize("kernel", ["tc", "s2d"])
@pytest.mark.parametrize("hw", [(64, 96), (480, 640), (70, 90)])
def test_stem_and_gate_match_oracle(hw, kernel):
    """fp32 stem + global gate vs the CPU oracle; gate logits to 1e-4 relative, hard decisions
    exact wherever the oracle's top-2 logit margin exceeds 1e-4 of the logit scale."""
    from dynmm_b200 import ops
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    h, w = hw
    b = 3 if h > 100 else 5
    cfg = fo.FusionConfig(height=h, width=w)
    sd = fo.make_state_dict(cfg, 0, 40.0)
    rgb, depth = sample_inputs(11, b, h, w)
    c = fo._Ctx(sd, False, "relu")
    with torch.no_grad():
        r = fo.encoder_first_conv(c, "encoder_rgb", rgb)
        d = fo.encoder_first_conv(c, "encoder_depth", depth)
        r = F.max_pool2d(r + d, 3, 2, 1)
        d = F.max_pool2d(d, 3, 2, 1)
        logits_ref = fo.global_gate_logits(c, r, d) if min(r.shape[2:]) >= 13 else None
    wr, sr, br = _stem_weights(sd, "encoder_rgb")
    wd, sdp, bd = _stem_weights(sd, "encoder_depth")
    if kernel == "s2d":      # TMA-gathered im2col (dynmm_stem_s2d_fwd)
        packed = ops.stem_s2d_pack_weights(wr, wd)
        r32, d32, r16, d16 = ops.stem_s2d(rgb.cuda(), depth.cuda(), packed, sr, br, sdp, bd)
        # BN vectors as a kernel parameter (the engine's path): the same values through the constant bank, bit for bit
        bn_host = ops.stem_s2d_bn_host(sr, br, sdp, bd)
        r32c, d32c, r16c, d16c = ops.stem_s2d(rgb.cuda(), depth.cuda(), packed, sr, br, sdp, bd, bn_host=bn_host)
        assert torch.equal(r32c, r32) and torch.equal(d32c, d32) and torch.equal(r16c, r16) and torch.equal(d16c, d16)
        only32 = ops.stem_s2d(rgb.cuda(), depth.cuda(), packed, sr, br, sdp, bd, bn_host=bn_host, want_bf16=False)
        assert only32[2] is None and torch.equal(only32[0], r32) and torch.equal(only32[1], d32)
        # [hi | lo] halves for the fp32-grade engine mode, written by the stem itself
        _, _, rs, ds = ops.stem_s2d(rgb.cuda(), depth.cuda(), packed, sr, br, sdp, bd, bn_host=bn_host, split=True)
        assert torch.equal(rs, ops.split_from_f32(r32)) and torch.equal(ds, ops.split_from_f32(d32))
    else:
        r32, d32, r16, d16 = ops.stem(rgb.cuda(), depth.cuda(), wr, sr, br, wd, sdp, bd)
    torch.cuda.synchronize()
    np.testing.assert_allclose(r32.permute(0, 3, 1, 2).cpu().numpy(), r.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(d32.permute(0, 3, 1, 2).cpu().numpy(), d.numpy(), rtol=1e-4, atol=1e-4)
    assert torch.equal(r16, r32.to(torch.bfloat16)) and torch.equal(d16, d32.to(torch.bfloat16))
    if logits_ref is None:
        return
    from dynmm_b200.fusion.engine import pack_gate
    gw = {k: v.cuda() for k, v in pack_gate(sd).items()}
    logits = ops.global_gate_logits(r32, d32, gw["w1"], gw["s1"], gw["b1"], gw["w2"], gw["s2"], gw["b2"], gw["wfc"])
    torch.cuda.synchronize()
    scale = logits_ref.abs().max().item()
    np.testing.assert_allclose(logits.cpu().numpy(), logits_ref.numpy(), rtol=1e-4, atol=1e-4 * scale)
    top2 = logits_ref.topk(2, 1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-4 * scale
    y, ys, idx = ops.diffsoftmax_fwd(logits, 1.0, True)
    ref_w = fo.diff_softmax(logits_ref, 1.0, True, 1)
    assert torch.equal(y.cpu()[decided], ref_w[decided]), "hard gate decisions differ from the fp32 oracle"
    assert decided.sum().item() >= 1


def test_diffsoftmax_against_golden(golden_dir):
    import os
    from dynmm_b200 import ops
    gold = np.load(os.path.join(golden_dir, "diffsoftmax.npz"))
    for name in ("b16x5", "b128x2", "ties"):
        logits = torch.from_numpy(gold[name + "_logits"]).cuda()
        up = torch.from_numpy(gold[name + "_upstream"]).cuda()
        for tau in (1.0, 0.5, 1e-3):
            for hard in (False, True):
                tag = f"{name}_tau{tau}_{'hard' if hard else 'soft'}"
                y, ys, idx = ops.diffsoftmax_fwd(logits, tau, hard)
                grad = ops.diffsoftmax_bwd(up, ys, tau)
                if hard:
                    np.testing.assert_array_equal(y.cpu().numpy(), gold[tag + "_y"])     # bit-exact one-hot
                else:
                    np.testing.assert_allclose(y.cpu().numpy(), gold[tag + "_y"], rtol=2e-6, atol=1e-7)
                g_ref = gold[tag + "_grad"]
                np.testing.assert_allclose(grad.cpu().numpy(), g_ref, rtol=2e-4, atol=2e-6 * max(1.0, np.abs(g_ref).max()))


def test_gate_plan():
    from dynmm_b200 import ops
    branches = torch.tensor([3, 0, 4, 1, 0, 2, 4, 3])
    w = torch.eye(5)[branches].cuda()
    hist = torch.zeros(5, dtype=torch.int64, device="cuda")
    plan = ops.gate_plan(w, hist=hist)
    torch.cuda.synchronize()
    g = plan.g.cpu()
    for s in range(4):
        assert torch.equal(g[s], (branches >= s + 1).float())
    assert plan.count.cpu().tolist() == [int((branches >= s).sum()) for s in (1, 2, 3, 4)]
    perm = plan.perm.cpu().long()
    need = branches[perm]
    assert torch.equal(need, need.sort(descending=True, stable=True).values)
    assert sorted(perm.tolist()) == list(range(8))
    assert torch.equal(plan.slot.cpu().long()[perm], torch.arange(8))
    assert hist.cpu().tolist() == [2, 1, 1, 2, 2]
    # soft weights: every stage mixes depth for every sample
    ws = torch.softmax(torch.randn(4, 5), 1).cuda()
    plan = ops.gate_plan(ws)
    assert plan.count.cpu().tolist() == [4, 4, 4, 4]
    ref = torch.stack([1 - ws[:, 0], 1 - (ws[:, 0] + ws[:, 1]), 1 - (ws[:, 0] + ws[:, 1] + ws[:, 2]), ws[:, 4]])
    np.testing.assert_allclose(plan.g.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_elementwise_ops():
    from dynmm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(5, 6, 7, 64, device=dev, generator=g).to(torch.bfloat16)
    b = torch.randn(5, 6, 7, 64, device=dev, generator=g).to(torch.bfloat16)
    gate = torch.tensor([1.0, 0.0, 0.5, 1.0, 0.0], device=dev)
    slot = torch.tensor([2, 0, 4, 1, 3], dtype=torch.int32, device=dev)
    out = ops.gated_add(a, b, gate, slot)
    ref = a.float() + gate.view(-1, 1, 1, 1) * b.float()[slot.long()]
    _bf16_close(out, ref, "gated_add")
    # fp32 training-path op and its gradients against autograd
    a32 = torch.randn(4, 64, 9, 11, device=dev, generator=g)
    a32 = a32[..., :8].contiguous()
    b32 = torch.randn_like(a32)
    gt = torch.rand(4, device=dev, generator=g)
    out = ops.gated_add_f32_fwd(a32, b32, gt)
    br, gr = b32.clone().requires_grad_(True), gt.clone().requires_grad_(True)
    ref = a32 + gr.view(-1, 1, 1, 1) * br
    np.testing.assert_allclose(out.cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-6, atol=1e-6)
    up = torch.randn_like(a32)
    ref.backward(up)
    gb, gg = ops.gated_add_f32_bwd(up, b32, gt)
    np.testing.assert_allclose(gb.cpu().numpy(), br.grad.cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gg.cpu().numpy(), gr.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    # layout round trip
    x = torch.randn(2, 40, 9, 13, device=dev, generator=g)
    nhwc = ops.nchw_f32_to_nhwc_bf16(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    back = ops.nhwc_bf16_to_nchw_f32(nhwc)
    assert torch.equal(back, x.to(torch.bfloat16).float())
    # learned upsampling: nearest x2 + depthwise 3x3 (+ skip)
    xin = torch.randn(2, 7, 9, 40, device=dev, generator=g).to(torch.bfloat16)
    wdw = torch.randn(40, 1, 3, 3, device=dev, generator=g) * 0.3
    bias = torch.randn(40, device=dev, generator=g) * 0.1
    skip = torch.randn(2, 14, 18, 40, device=dev, generator=g).to(torch.bfloat16)
    up_ref = F.conv2d(F.interpolate(xin.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest"), wdw, bias, 1, 1,
                      1, 40)
    got = ops.upsample2x_dw3x3(xin, wdw.view(40, 9).t().contiguous(), bias, skip)
    _bf16_close(got, up_ref.permute(0, 2, 3, 1) + skip.float(), "upsample+skip")
    got32 = ops.upsample2x_dw3x3(xin, wdw.view(40, 9).t().contiguous(), bias, to_nchw_f32=True)
    np.testing.assert_allclose(got32.cpu().numpy(), up_ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
    lab = torch.empty(2, 14, 18, dtype=torch.uint8, device=dev)
    got_l = ops.upsample2x_dw3x3(xin, wdw.view(40, 9).t().contiguous(), bias, labels=lab)
    assert torch.equal(got_l, got32) and torch.equal(lab.long(), got32.argmax(1))        # fused arg-max
    lab2 = torch.zeros_like(lab)
    assert ops.upsample2x_dw3x3(xin, wdw.view(40, 9).t().contiguous(), bias, labels=lab2, want_logits=False) is None
    assert torch.equal(lab2, lab)
    # pyramid pooling helpers
    feat = torch.randn(2, 15, 20, 64, device=dev, generator=g).to(torch.bfloat16)
    for bins in (1, 5):
        p = ops.adaptive_avgpool(feat, bins)
        ref = F.adaptive_avg_pool2d(feat.float().permute(0, 3, 1, 2), bins).permute(0, 2, 3, 1)
        _bf16_close(p, ref, f"avgpool{bins}")
        dst = torch.zeros(2, 15, 20, 128, dtype=torch.bfloat16, device=dev)
        ops.nearest_resize_into(p, dst, 64)
        ref_up = F.interpolate(p.float().permute(0, 3, 1, 2), (15, 20), mode="nearest").permute(0, 2, 3, 1)
        assert torch.equal(dst[..., 64:].float(), ref_up) and (dst[..., :64] == 0).all()


def test_softgate_mix_and_compaction():
    from dynmm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(2)
    b, c = 128, 23
    p0 = torch.randn(b, c, device=dev, generator=g)
    p1 = torch.randn(b, c, device=dev, generator=g)
    w = torch.softmax(torch.randn(b, 2, device=dev, generator=g), 1)
    out = ops.softgate_mix_fwd([p0, p1], w)
    ref = w[:, 0:1] * p0 + w[:, 1:2] * p1
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6, atol=1e-6)
    up = torch.randn(b, c, device=dev, generator=g)
    grads, gw = ops.softgate_mix_bwd(up, [p0, p1], w, [True, True])
    np.testing.assert_allclose(grads[0].cpu().numpy(), (w[:, 0:1] * up).cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gw[:, 1].cpu().numpy(), (up * p1).sum(1).cpu().numpy(), rtol=1e-4, atol=1e-4)
    # hard gate: expert 1 evaluated on the compacted rows only
    hard = torch.eye(2, device=dev)[torch.randint(0, 2, (b,), device=dev, generator=g)]
    idx, inv, cnt = ops.compact_rows(hard, 1)
    k = int(cnt.item())
    assert k == int(hard[:, 1].sum().item())
    assert torch.equal(idx[:k].long(), hard[:, 1].nonzero().flatten())
    p1_small = torch.full((b, c), float("nan"), device=dev)
    p1_small[:k] = p1[idx[:k].long()]
    out = ops.softgate_mix_fwd([p0, p1_small], hard, rows=[None, inv])
    ref = hard[:, 0:1] * p0 + hard[:, 1:2] * p1
    assert torch.equal(out, ref)


def test_se_fusion_kernels():
    """Squeeze (deterministic GAP), excite (MLP + sigmoid) and the gated SE blend vs PyTorch."""
    from dynmm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(4)
    n, h, w, c = 5, 15, 20, 128
    rgb = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    depth = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    w1 = torch.randn(c // 16, c, device=dev, generator=g) * 0.2
    b1 = torch.randn(c // 16, device=dev, generator=g) * 0.1
    w2 = torch.randn(c, c // 16, device=dev, generator=g) * 0.2
    b2 = torch.randn(c, device=dev, generator=g) * 0.1
    part = ops.gap_partial(rgb)
    np.testing.assert_allclose(part.sum(1).cpu().numpy() / (h * w), rgb.float().mean((1, 2)).cpu().numpy(),
                               rtol=1e-4, atol=1e-5)
    assert torch.equal(part, ops.gap_partial(rgb))                       # deterministic
    sig = ops.se_mlp(part, 1.0 / (h * w), w1, b1, w2, b2)
    mean = rgb.float().mean((1, 2))
    ref_sig = torch.sigmoid(F.relu(mean @ w1.t() + b1) @ w2.t() + b2)
    np.testing.assert_allclose(sig.cpu().numpy(), ref_sig.cpu().numpy(), rtol=1e-4, atol=1e-5)
    sig_d = torch.rand(n, c, device=dev, generator=g)
    gate = torch.tensor([1.0, 0.0, 0.3, 1.0, 0.0], device=dev)
    slot = torch.tensor([1, 4, 0, 2, 3], dtype=torch.int32, device=dev)
    depth_nan = depth.clone()
    depth_nan[4] = float("nan")                                          # slot of a gated-off sample: never read
    depth_nan[3] = float("nan")
    sig_d_nan = sig_d.clone()
    sig_d_nan[4] = float("nan")
    sig_d_nan[3] = float("nan")
    out = ops.se_gated_fuse(rgb, depth_nan, sig, sig_d_nan, gate, slot)
    gg = gate.view(-1, 1, 1, 1)
    d_sel = torch.nan_to_num(depth_nan.float()[slot.long()])
    sd_sel = torch.nan_to_num(sig_d_nan[slot.long()])
    ref = rgb.float() * (1 - gg + gg * sig.view(n, 1, 1, c)) + gg * sd_sel.view(n, 1, 1, c) * d_sel
    _bf16_close(out, ref, "se_gated_fuse")
    assert torch.isfinite(out.float()).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,has_bias,has_bn", [((64, 64, 3, 1), True, False), ((128, 64, 1, 3), True, True),
                                                   ((40, 128, 3, 3), False, False), ((256, 128, 1, 1), False, True),
                                                   ((24, 16, 3, 3), True, True)])
def test_fold_pack_conv_is_bit_identical_to_the_pytorch_expression(shape, has_bias, has_bn):
    """dynmm_fold_pack_conv (engine build, one launch per convolution) == fold_bn + scale + pack_conv_weight done with
    fp32 PyTorch ops (what the oracle folds: conv + eval BatchNorm, resnet.py:124-147), bit for bit."""
    from dynmm_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    co = shape[0]
    w = (torch.randn(shape, generator=g) * 0.1).cuda()
    bias = torch.randn(co, generator=g).cuda() if has_bias else None
    bn = None
    if has_bn:
        bn = [(0.5 + torch.rand(co, generator=g)).cuda(), torch.randn(co, generator=g).cuda(),
              (torch.randn(co, generator=g) * 0.2).cuda(), (0.3 + torch.rand(co, generator=g)).cuda()]
    packed, shift = ops.fold_pack_conv(w, bias, bn, 1e-3)
    if has_bn:
        scale, ref_shift = ops.fold_bn(*bn, 1e-3, bias)
        ref_packed = ops.pack_conv_weight(w * scale.view(-1, 1, 1, 1))
        s2, b2 = ops.fold_bn_cuda(*bn, 1e-3, bias)
        assert torch.equal(s2, scale) and torch.equal(b2, ref_shift)
    else:
        ref_packed, ref_shift = ops.pack_conv_weight(w), bias
    assert torch.equal(packed.view(torch.int16), ref_packed.view(torch.int16))
    if ref_shift is None:
        assert shift is None
    else:
        assert torch.equal(shift, ref_shift)


@pytest.mark.gpu
def test_permute3d_matches_torch():
    from dynmm_b200 import ops
    x = torch.randn(7, 5, 9).cuda()
    for perm in [(2, 1, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (2, 0, 1), (1, 2, 0)]:
        assert torch.equal(ops.permute3d(x, perm), x.permute(*perm).contiguous())


DUAL_CASES = [
    # name, n, h, w, cin, cout, kh, kw, stride, pad  (streamed-weight layers: C >= 256)
    ("s3_1x3_c256", 8, 30, 40, 256, 256, 1, 3, (1, 1), (0, 1)),
    ("s3_3x1_c256", 8, 30, 40, 256, 256, 3, 1, (1, 1), (1, 0)),
    ("s4_1x3_c512", 8, 15, 20, 512, 512, 1, 3, (1, 1), (0, 1)),
    ("s4_3x1_c512_n5", 5, 15, 20, 512, 512, 3, 1, (1, 1), (1, 0)),          # odd number of pixel tiles
    ("s3_3x1_s2_128to256", 3, 60, 80, 128, 256, 3, 1, (2, 1), (1, 0)),      # strided: per-tap loads
    ("s4_1x1_s2_ds", 8, 30, 40, 256, 512, 1, 1, (2, 2), (0, 0)),
    ("ppm_1x1_c768", 8, 15, 20, 768, 128, 1, 1, (1, 1), (0, 0)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", DUAL_CASES, ids=[c[0] for c in DUAL_CASES])
def test_conv_dual_units_are_bit_identical(case):
    """Dual-M work units (two pixel tiles share every streamed weight tile; DYNMM_CONV_FORCE_DUAL) give the bits of the
    one-tile units (DYNMM_CONV_NO_DUAL): same UMMA sequence per tile.  Full epilogue (shift, residual through res_map,
    ReLU, gated add), sample indirection and every device-side count, and the fp32 reference on top."""
    from dynmm_b200 import ops
    name, n, h, w, cin, cout, kh, kw, stride, pad = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(len(name) + n)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(cout, cin, kh, kw, device=dev, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5
    packed = ops.pack_conv_weight(wt)
    ho = (h + 2 * pad[0] - kh) // stride[0] + 1
    wo = (w + 2 * pad[1] - kw) // stride[1] + 1
    shift = torch.randn(cout, device=dev, generator=g) * 0.1
    res = torch.randn(n, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
    gated = torch.randn(n, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
    gate = torch.rand(n, device=dev, generator=g)
    gate[0] = 0.0
    slot = torch.randperm(n, device=dev, generator=g).to(torch.int32)
    perm = torch.randperm(n, device=dev, generator=g).to(torch.int32)
    base = dict(c_out=cout, kh=kh, kw=kw, stride=stride, pad=pad, shift=shift, relu=True)
    variants = [dict(), dict(residual=res), dict(residual=res, gated=gated, gate=gate, gated_slot=slot),
                dict(in_map=perm, residual=res, res_map=perm)]
    for kw_ in variants:
        a = ops.conv(x, packed, dual=False, **base, **kw_)
        b = ops.conv(x, packed, dual=True, **base, **kw_)
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (name, sorted(kw_))
    ref = _ref_conv(x, wt, stride, pad, None, shift, res, True, gated, gate, slot)
    _bf16_close(ops.conv(x, packed, dual=True, residual=res, gated=gated, gate=gate, gated_slot=slot, **base), ref, name)
    for cnt in range(0, n + 1):
        count = torch.tensor([cnt], dtype=torch.int32, device=dev)
        outs = []
        for dual in (False, True):
            out = torch.full((n, ho, wo, cout), 7.0, dtype=torch.bfloat16, device=dev)
            ops.conv(x, packed, dual=dual, count=count, in_map=perm, residual=res, res_map=perm, out=out, **base)
            outs.append(out)
        torch.cuda.synchronize()
        assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16)), (name, cnt)
        assert (outs[1][cnt:].float() == 7.0).all()


def _flag_chain(x, layers, pool):
    """A NonBottleneck1D-like chain through ops.conv; with a pool the launches publish / consume tile flags."""
    from dynmm_b200 import ops
    ops.FLAG_POOL = pool
    try:
        if pool is not None:
            pool.reset()
        outs = []
        y, block_in = x, x
        for i, (packed, shift, geom, use_res) in enumerate(layers):
            cout, kh, kw, stride, pad = geom
            res = block_in if (use_res and block_in.shape[3] == cout and stride == (1, 1) and
                               block_in.shape[1:3] == y.shape[1:3]) else None
            y = ops.conv(y, packed, c_out=cout, kh=kh, kw=kw, stride=stride, pad=pad, shift=shift, relu=True,
                         residual=res, residual_settled=res is not None and not hasattr(res, "_dynmm_flags"))
            outs.append(y)
            if use_res:
                block_in = y
        return outs
    finally:
        ops.FLAG_POOL = None


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 60, 80, 128), (8, 30, 40, 256), (6, 15, 20, 512), (3, 24, 40, 64)],
                         ids=["s2_c128", "s3_c256", "s4_c512", "c64"])
def test_tile_flags_chain_is_bit_identical_under_graph_replay(shape):
    """Layer-to-layer overlap through tile-completion flags (dynmm_tile_flags): a chain of 3x1 / 1x3 / strided
    convolutions whose launches wait on their producers' flags instead of on the previous kernel gives exactly the
    bits of the stream-ordered chain -- eager, and replayed 30 times from a CUDA graph (where consecutive launches
    really overlap through programmatic dependent launch)."""
    from dynmm_b200 import ops
    n, h, w, c = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(c + h)
    x = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
    geoms = [(c, 3, 1, (1, 1), (1, 0), False), (c, 1, 3, (1, 1), (0, 1), False),
             (c, 3, 1, (1, 1), (1, 0), False), (c, 1, 3, (1, 1), (0, 1), True),
             (c, 3, 1, (1, 1), (1, 0), False), (c, 1, 3, (1, 1), (0, 1), True),
             (2 * c, 3, 1, (2, 1), (1, 0), False), (2 * c, 1, 3, (1, 2), (0, 1), False),
             (2 * c, 3, 3, (1, 1), (1, 1), False), (2 * c, 1, 1, (1, 1), (0, 0), True)]
    layers = []
    cin = c
    for cout, kh, kw, stride, pad, use_res in geoms:
        wt = torch.randn(cout, cin, kh, kw, device=dev, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5
        layers.append((ops.pack_conv_weight(wt), torch.randn(cout, device=dev, generator=g) * 0.1,
                       (cout, kh, kw, stride, pad), use_res))
        cin = cout
    ref = _flag_chain(x, layers, None)
    pool = ops.TileFlagPool(dev)
    got = _flag_chain(x, layers, pool)
    torch.cuda.synchronize()
    assert any(hasattr(t, "_dynmm_flags") for t in got), "no launch published flags"
    for i, (a, b) in enumerate(zip(ref, got)):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), f"eager, layer {i}"
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _flag_chain(x, layers, pool)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = _flag_chain(x, layers, pool)
    for rep in range(30):
        for t in outs:
            t.fill_(3.0)
        graph.replay()
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(ref, outs)):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16)), f"graph replay {rep}, layer {i}"


MERGE_CASES = [
    # name, n_a, n_b, h, w, cin, cout, kh, kw, stride, pad
    ("s2_3x1_c128", 8, 8, 60, 80, 128, 128, 3, 1, (1, 1), (1, 0)),
    ("s3_1x3_c256", 8, 8, 30, 40, 256, 256, 1, 3, (1, 1), (0, 1)),
    ("s4_3x1_c512", 8, 8, 15, 20, 512, 512, 3, 1, (1, 1), (1, 0)),
    ("s3_3x1_s2_128to256", 8, 8, 60, 80, 128, 256, 3, 1, (2, 1), (1, 0)),
    ("s3_1x1_s2_ds", 8, 8, 60, 80, 128, 256, 1, 1, (2, 2), (0, 0)),
    ("small_c64_3x3", 5, 5, 12, 20, 64, 64, 3, 3, (1, 1), (1, 1)),
    ("s3_1x3_c256_n8_n6", 8, 6, 30, 40, 256, 256, 1, 3, (1, 1), (0, 1)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_merged_launch_is_bit_identical(case):
    """dynmm_conv_igemm_fwd2 (the same layer of the RGB and of the depth encoder in ONE launch, ops.ConvMerge): each
    job's output has the bits of its own dynmm_conv_igemm_fwd launch -- for every device-side count of the second
    job (0 .. n: gated-off depth samples), with residuals, and with nothing written beyond the count."""
    from dynmm_b200 import ops
    name, na, nb, h, w, cin, cout, kh, kw, stride, pad = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(len(name) * 7 + na)
    xa = torch.randn(na, h, w, cin, device=dev, generator=g).to(torch.bfloat16)
    xb = torch.randn(nb, h, w, cin, device=dev, generator=g).to(torch.bfloat16)
    mk = lambda: ops.pack_conv_weight(torch.randn(cout, cin, kh, kw, device=dev, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5)
    wa, wb = mk(), mk()
    sa, sb = torch.randn(cout, device=dev, generator=g) * 0.1, torch.randn(cout, device=dev, generator=g) * 0.1
    ho = (h + 2 * pad[0] - kh) // stride[0] + 1
    wo = (w + 2 * pad[1] - kw) // stride[1] + 1
    ra = torch.randn(na, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
    rb = torch.randn(nb, ho, wo, cout, device=dev, generator=g).to(torch.bfloat16)
    base = dict(c_out=cout, kh=kh, kw=kw, stride=stride, pad=pad, relu=True)
    for with_res in (False, True):
        for cnt in sorted({0, 1, nb // 2, nb - 1, nb}):
            count = torch.tensor([cnt], dtype=torch.int32, device=dev)
            kw_a = dict(shift=sa, residual=ra if with_res else None, **base)
            kw_b = dict(shift=sb, residual=rb if with_res else None, count=count, **base)
            ref_a = ops.conv(xa, wa, **kw_a)
            ref_b = torch.full((nb, ho, wo, cout), 7.0, dtype=torch.bfloat16, device=dev)
            ops.conv(xb, wb, out=ref_b, **kw_b)
            got_b = torch.full((nb, ho, wo, cout), 7.0, dtype=torch.bfloat16, device=dev)
            with ops.ConvMerge() as m:
                got_a = ops.conv(xa, wa, **kw_a)
                ops.conv(xb, wb, out=got_b, **kw_b)
            torch.cuda.synchronize()
            assert m.merged, f"{name}: the two convolutions were not merged"
            assert torch.equal(got_a.view(torch.int16), ref_a.view(torch.int16)), (name, with_res, cnt, "job a")
            assert torch.equal(got_b.view(torch.int16), ref_b.view(torch.int16)), (name, with_res, cnt, "job b")
            assert (got_b[cnt:].float() == 7.0).all()
    # different geometry: not merged, still correct
    with ops.ConvMerge() as m:
        y1 = ops.conv(xa, wa, shift=sa, **base)
        y2 = ops.conv(xb[:, : h // 2 + 1].contiguous(), wb, shift=sb, **base)
    torch.cuda.synchronize()
    assert not m.merged
    assert torch.equal(y1.view(torch.int16), ops.conv(xa, wa, shift=sa, **base).view(torch.int16))
