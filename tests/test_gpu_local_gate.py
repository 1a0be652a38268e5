"""Local-gate dynamic ESANet on the GPU.

* The Gumbel gate op (noise drawn like ``F.gumbel_softmax`` draws it + the ``diff_softmax`` CUDA kernel) makes the
  SAME decisions as ``F.gumbel_softmax`` on the same device and seed, soft values to 1e-6, and has its gradient.
* With the random policy (CPU ``torch.randint``, rgb_depth_fusion.py:36-40) nothing depends on the device generator, so
  the reference's own vectors (tests/golden/local_gate_*.npz) apply: decisions bit-exact, logits within the fp32 /
  stated bf16 tolerance -- the fp32 graph uses the ``gated_add`` kernel for the blends, ``train_precision='bf16'``
  additionally runs every stage convolution on the tcgen05 kernels.
"""
import os

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


@pytest.mark.parametrize("hard", [True, False])
def test_gumbel_gate_equals_torch_gumbel_softmax_on_device(hard):
    from dynmm_b200.fusion.local_gate import gumbel_softmax
    g = torch.Generator().manual_seed(9)
    logits = (torch.rand(64, 2, generator=g) / 0.7).cuda()
    torch.manual_seed(123)
    ref = F.gumbel_softmax(logits, tau=1, hard=hard)
    torch.manual_seed(123)
    x = logits.clone().requires_grad_(True)
    got = gumbel_softmax(x, hard)
    if hard:
        assert torch.equal(got.detach().argmax(1), ref.argmax(1))            # decisions bit-exact
        torch.testing.assert_close(got.detach(), ref, rtol=0, atol=2e-7)     # one-hot up to (1 - s) + s rounding
    else:
        torch.testing.assert_close(got.detach(), ref, rtol=1e-5, atol=1e-6)
    # straight-through / soft gradient = Jacobian of the softmax of the noisy logits
    torch.manual_seed(123)
    y = logits.clone().requires_grad_(True)
    up = torch.randn(64, 2, generator=g).cuda()
    (F.gumbel_softmax(y, tau=1, hard=hard) * up).sum().backward()
    (got * up).sum().backward()
    torch.testing.assert_close(x.grad, y.grad, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["local_gate_r18_basic_64x64", "local_gate_r34_nbt1d_64x96"])
def test_random_policy_matches_reference_vectors(name, golden_dir):
    from dynmm_b200.fusion import SkipESANet
    from oracle.make_golden_local import CASES, MODES, apply_mode, sample_inputs, seeded_state
    kw, seed, b = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    model = SkipESANet(pretrained_on_imagenet=False, **kw)
    model.load_state_dict(seeded_state(model.state_dict(), seed), strict=True)
    model = model.cuda().eval()
    model.use_engine = False                  # this test is about the differentiable graph
    rgb, depth = (t.cuda() for t in sample_inputs(seed + 100, b, kw["height"], kw["width"]))
    tag, rule, attrs, test, fseed = next(m for m in MODES if m[0] == "random")
    ref = torch.from_numpy(gold[f"{tag}_out"])
    for precision, tol in (("fp32", 2e-3), ("bf16", 2e-2)):
        model.train_precision = precision
        apply_mode(model, rule, attrs)
        model.start_weight()
        torch.manual_seed(fseed)
        with torch.no_grad():
            out = model(rgb, depth, test)
        model._flush_weights()
        for i in range(4):
            np.testing.assert_array_equal(model.weight_list[i].numpy(), gold[f"{tag}_weight{i}"])
        model.end_weight()
        got = out[:, :, ::4, ::4].float().cpu()
        err = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
        assert err <= tol, f"{name}/{precision}: relative L2 error {err:.5f}"
    # static rules (no gate in the data path): 1111 = plain ESANet-style add fusion
    tag, rule, attrs, test, fseed = next(m for m in MODES if m[0] == "static1111")
    model.train_precision = "fp32"
    apply_mode(model, rule, attrs)
    torch.manual_seed(fseed)
    with torch.no_grad():
        out = model(rgb, depth, test)
    ref = torch.from_numpy(gold[f"{tag}_out"])
    err = ((out[:, :, ::4, ::4].cpu().double() - ref.double()).norm() / ref.double().norm()).item()
    assert err <= 2e-3, err


def test_hard_chain_is_monotone_on_device():
    """Chained hard gates (prev_weight, model_skip_mod.py:259,277,295): once a site is closed every later one is."""
    from dynmm_b200.fusion import SkipESANet
    from oracle.make_golden_local import CASES, apply_mode, sample_inputs, seeded_state
    kw, seed, b = CASES["local_gate_r18_basic_64x64"]
    model = SkipESANet(pretrained_on_imagenet=False, **kw)
    model.load_state_dict(seeded_state(model.state_dict(), seed), strict=True)
    model = model.cuda().eval()
    model.use_engine = False
    rgb, depth = (t.cuda() for t in sample_inputs(seed + 100, 8, kw["height"], kw["width"]))
    apply_mode(model, [2, 2, 2, 2], dict(hard_gate=True))
    model.start_weight()
    with torch.no_grad():
        for s in range(4):
            torch.manual_seed(s)
            out = model(rgb, depth, True)
            assert torch.isfinite(out).all()
    model._flush_weights()
    w = torch.stack([model.weight_list[i][:, 1] for i in range(4)])            # [site, samples]
    assert (w - w.round()).abs().max().item() <= 1e-6            # one-hot up to the (1 - s) + s rounding of the ST form
    w = w.round()
    assert (w[1:] <= w[:-1]).all()
    assert 0 < w.sum() < w.numel()
    model.end_weight()


def _r34_model(case="local_gate_r34_nbt1d_64x96"):
    from dynmm_b200.fusion import SkipESANet
    from oracle.make_golden_local import CASES, seeded_state
    kw, seed, b = CASES[case]
    model = SkipESANet(pretrained_on_imagenet=False, **kw)
    model.load_state_dict(seeded_state(model.state_dict(), seed), strict=True)
    return model.cuda().eval(), kw, seed, b


@pytest.mark.parametrize("precision", ["bf16", "f32x3"])
@pytest.mark.parametrize("case", ["local_gate_r34_nbt1d_64x96", "local_gate_r18_basic_64x64"])
@pytest.mark.parametrize("tag", ["random", "static1111"])
def test_engine_matches_reference_vectors(tag, case, precision, golden_dir):
    """Eval mode on CUDA runs FusionEngine.forward_local (per-stage skipping; bf16 kernels within 2e-2, the fp32-grade
    mode within 1e-3 of the reference's fp32 outputs).  Modes whose decisions do not depend on the device generator
    have reference vectors: random policy (CPU randint) and the static rule."""
    from oracle.make_golden_local import MODES, apply_mode, sample_inputs
    # the second case: ResNet-18 BasicBlock encoders, bilinear up-sampling, 37 classes (model_skip_mod.py defaults)
    model, kw, seed, b = _r34_model(case)
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    rgb, depth = (t.cuda() for t in sample_inputs(seed + 100, b, kw["height"], kw["width"]))
    _, rule, attrs, test, fseed = next(m for m in MODES if m[0] == tag)
    apply_mode(model, rule, attrs)
    assert model.use_engine
    model.engine_precision = precision
    model.start_weight()
    torch.manual_seed(fseed)
    with torch.no_grad():
        out = model(rgb, depth, test)
    assert model._engine is not None and getattr(model, "_engine_unsupported", None) is None
    assert model._engine.split == (precision == "f32x3")
    model._flush_weights()
    if tag == "random":
        for i in range(4):
            np.testing.assert_array_equal(model.weight_list[i].numpy(), gold[f"{tag}_weight{i}"])
        # chained one-hot decisions: stage s + 1 of the depth encoder only ran for samples still open after site s
        w1 = np.stack([gold[f"{tag}_weight{i}"][:, 1] for i in range(4)])
        counts = [int(c.item()) for c in model.last_counts]
        assert counts[0] == int((w1[0] != 0).sum())
        assert all(counts[s] <= counts[s - 1] for s in range(1, 4))
    model.end_weight()
    ref = torch.from_numpy(gold[f"{tag}_out"])
    got = out[:, :, ::4, ::4].float().cpu()
    err = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
    assert err <= (1e-3 if precision == "f32x3" else 2e-2), f"{tag}/{precision}: relative L2 error {err:.5f}"


@pytest.mark.parametrize("rule,attrs,test", [
    ([2, 2, 2, 2], dict(), True),                      # hard Gumbel gates, chained
    ([2, 2, 2, 2], dict(hard_gate=True, ini_stage=True), False),
    ([1, 1, 2, 2], dict(hard_gate=True), True),
    ([0, 1, 2, 0], dict(), True),
    ([2, 2, 2, 2], dict(), False),                     # soft gates: nothing can be skipped
])
def test_engine_matches_module_graph_on_device(rule, attrs, test):
    """Same seed, same device generator: the engine draws the Gumbel noise in the order the module graph does, so the
    decisions agree and the logits differ by the bf16 tolerance only; skipped depth work shows in the stage counts."""
    from oracle.make_golden_local import apply_mode, sample_inputs
    model, kw, seed, _ = _r34_model()
    rgb, depth = (t.cuda() for t in sample_inputs(seed + 7, 6, kw["height"], kw["width"]))
    apply_mode(model, rule, attrs)
    res = {}
    for use_engine in (False, True):
        model.use_engine = use_engine
        model.start_weight()
        torch.manual_seed(77)
        with torch.no_grad():
            out = model(rgb, depth, test)
        model._flush_weights()
        res[use_engine] = (out.float().cpu(), [w.clone() for w in model.weight_list])
        model.end_weight()
    hard = test or attrs.get("hard_gate", False)
    for i in range(4):
        if rule[i] != 2:
            continue      # the rule ignores this site's gate: the engine draws its noise (generator order) on dummy features
        a, b = res[False][1][i], res[True][1][i]
        if hard:
            assert torch.equal(a.round(), b.round()), f"site {i}: decisions differ"
        else:
            torch.testing.assert_close(a, b, rtol=0, atol=2e-2)
    ref, got = res[False][0], res[True][0]
    err = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
    assert err <= 2e-2, f"relative L2 error {err:.5f}"
    if hard and not attrs.get("ini_stage", False):
        counts = [int(c.item()) for c in model.last_counts]
        need = torch.ones(6, dtype=torch.bool)
        for s in range(4):
            # a sample needs depth stage s + 1 for this site's blend or, with an open chain, for a later one
            w = res[True][1][s][:, 1]
            later_add = any(r == 1 for r in rule[s + 1:])
            later_dyn = any(r == 2 for r in rule[s + 1:])
            g = {0: torch.zeros(6), 1: torch.ones(6), 2: w.round()}[rule[s]]
            assert counts[s] <= 6
            if not later_add and not later_dyn:
                assert counts[s] <= int((g != 0).sum())


def test_local_gate_engine_graph_replay_matches_eager_launches():
    """SkipESANet.use_cuda_graph: one captured graph (device-side slot planning, device generator) serves every decision;
    with a forced static rule the result is deterministic and must equal the eager launches bit for bit, with dynamic
    gates the replays stay finite and the chained counts monotone."""
    from oracle.make_golden_local import apply_mode, sample_inputs
    model, kw, seed, _ = _r34_model()
    rgb, depth = (t.cuda() for t in sample_inputs(seed + 3, 4, kw["height"], kw["width"]))
    apply_mode(model, [1, 0, 1, 1], dict())
    with torch.no_grad():
        eager = model(rgb, depth, True).clone()
        model.use_cuda_graph = True
        for _ in range(3):
            out = model(rgb, depth, True)
        assert torch.equal(out, eager)
        apply_mode(model, [2, 2, 2, 2], dict(hard_gate=True))
        for _ in range(4):
            out = model(rgb, depth, True)
            counts = [int(c.item()) for c in model.last_counts]
            assert torch.isfinite(out).all() and all(counts[s] <= counts[s - 1] for s in range(1, 4))
