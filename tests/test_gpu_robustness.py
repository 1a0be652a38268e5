"""BASELINE.json configs[4] on the GPU: the robustness sweep (dynmm_b200/fusion/robustness.py) drives the CUDA engine;
the per-sample hard gate decisions on the PERTURBED inputs must equal the fp32 oracle's on the very same tensors
(the noise is regenerated from the same seeds on the same device, eval.py:20-23, 91-102), for logits and for the
fused arg-max labels path, eager and CUDA-graph replay."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

MARGIN = 1e-3        # oracle top-2 logit margin (relative to the logit scale) below which a decision is not compared


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    yield


def _model_and_data():
    from dynmm_b200.fusion import SkipGateESANet
    from oracle import fusion_oracle as fo
    cfg = fo.FusionConfig(height=64, width=96)
    sd = fo.make_state_dict(cfg, 3, 40.0)
    model = SkipGateESANet(height=64, width=96)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    model.hard_gate = True
    g = torch.Generator().manual_seed(17)
    data = []
    for _ in range(5):
        gain = 0.25 + 1.5 * torch.rand(3, 2, generator=g)
        data.append(((torch.randn(3, 3, 64, 96, generator=g) * gain[:, :1].view(-1, 1, 1, 1)).cuda(),
                     (torch.randn(3, 1, 64, 96, generator=g) * gain[:, 1:].view(-1, 1, 1, 1)).cuda()))
    return model, sd, cfg, data


@pytest.mark.parametrize("labels_only,graph", [(False, False), (True, False), (False, True)])
def test_sweep_decisions_match_oracle_on_perturbed_inputs(labels_only, graph):
    from dynmm_b200.fusion import robustness as rb
    from oracle import fusion_oracle as fo
    model, sd, cfg, data = _model_and_data()
    model.use_cuda_graph = graph
    mode, noise, runs = 2, 1.0, 2
    # the perturbed inputs the sweep will see: same seeds, same device, same order of draws
    perturbed = []
    for r in range(runs):
        rb.set_seed(r)
        for image, depth in data:
            image, depth, which = rb.perturb(image, depth, mode, noise, random.random())
            perturbed.append((image.cpu(), depth.cpu(), which))
    got = []
    pt = rb.run_point(model, lambda r: data, mode, noise, runs, labels_only=labels_only,
                      on_batch=lambda r, i, pred, w: got.append((pred.clone(), w.clone())))
    torch.cuda.synchronize()
    assert pt.batches == len(perturbed) and pt.images == 3 * len(perturbed) and sum(pt.histogram) == pt.images
    assert pt.noised_batches == sum(w >= 0 for _, _, w in perturbed) and 0 < pt.noised_batches < pt.batches
    hist = [0] * 5
    compared = 0
    for (image, depth, _), (pred, w) in zip(perturbed, got):
        ref = fo.forward(sd, cfg, image, depth, hard_gate=True)
        top2 = ref["gate_logits"].topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > MARGIN * ref["gate_logits"].abs().max()
        ours, theirs = w.argmax(1).cpu(), ref["weight"].argmax(1)
        assert torch.equal(ours[safe], theirs[safe])
        assert torch.equal(w.cpu().sum(1), torch.ones(3)) and ((w == 0) | (w == 1)).all()     # one-hot
        compared += int(safe.sum())
        for k in ours.tolist():
            hist[k] += 1
        if labels_only:
            assert pred.dtype == torch.uint8 and tuple(pred.shape) == (3, 64, 96)
            with torch.no_grad():       # the fused arg-max equals the arg-max of the engine's own logits, bit for bit
                own = model(image.cuda(), depth.cuda(), True).argmax(1).to(torch.uint8)
            assert torch.equal(pred, own)
            agree = (pred.cpu().long() == ref["out"].argmax(1)).float().mean().item()
            assert agree >= 0.95, agree
        else:
            err = ((pred.cpu().double() - ref["out"].double()).norm() / ref["out"].double().norm()).item()
            assert err <= 2e-2, err
    assert compared >= 0.9 * pt.images
    assert hist == pt.histogram


def test_noise_changes_the_branch_distribution_only_through_the_inputs():
    """sigma = 0 under any mode is the clean sweep (x + 0 * ... == x up to the sign of zero): identical histograms."""
    from dynmm_b200.fusion import robustness as rb
    model, _, _, data = _model_and_data()
    clean = rb.run_point(model, lambda r: data, -1, 0.0, 1)
    zero = rb.run_point(model, lambda r: data, 1, 0.0, 1)
    assert clean.histogram == zero.histogram and zero.noised_batches > 0
    pts = rb.sweep(model, lambda r: data, noises=(0.0, 1.0), mode=1, num_runs=1)
    assert [p.noise for p in pts] == [0.0, 1.0] and all(p.saved_pct is not None and p.images_per_s > 0 for p in pts)
