"""Host logic of the robustness sweep (BASELINE.json configs[4]; dynmm_b200/fusion/robustness.py) on CPU: the
perturbation follows the reference's evaluation loop statement by statement (FusionDynMM/eval.py:20-23, 77-102), the
sweep's bookkeeping is consistent, and the 2-rank path (gloo) sums to the single-process result."""
import os
import random
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fusion_oracle as fo

warnings.filterwarnings("ignore")


def _reference_loop(batches, mode, noise, num_runs):
    """eval.py:77-102, verbatim control flow, returning the perturbed inputs of every batch of every run."""
    out = []
    for r in range(num_runs):
        random.seed(r)                      # set_seed(r), eval.py:20-23
        np.random.seed(r)
        torch.manual_seed(r)
        for image, depth in batches:
            rand_val = random.random()
            if mode == 0:
                if rand_val < 0.33:
                    image = image + noise * abs(image).mean() * torch.randn_like(image)
            elif mode == 1:
                if rand_val < 0.33:
                    depth = depth + noise * abs(depth).mean() * torch.randn_like(depth)
            elif mode == 2:
                if rand_val < 0.33:
                    image = image + noise * abs(image).mean() * torch.randn_like(image)
                elif rand_val < 0.66:
                    depth = depth + noise * abs(depth).mean() * torch.randn_like(depth)
            out.append((image, depth))
    return out


@pytest.mark.parametrize("mode", [-1, 0, 1, 2])
def test_perturbation_follows_the_reference_loop(mode):
    from dynmm_b200.fusion import robustness as rb
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(2, 3, 8, 12, generator=g), torch.randn(2, 1, 8, 12, generator=g) + 1.5) for _ in range(9)]
    ref = _reference_loop(batches, mode, 0.6, 2)
    got, whiches = [], []
    for r in range(2):
        rb.set_seed(r)
        for image, depth in batches:
            image, depth, which = rb.perturb(image, depth, mode, 0.6, random.random())
            got.append((image, depth))
            whiches.append(which)
    for (ri, rd), (gi, gd) in zip(ref, got):
        assert torch.equal(ri, gi) and torch.equal(rd, gd)
    if mode == -1:
        assert set(whiches) == {-1}
    else:
        assert any(w >= 0 for w in whiches) and any(w < 0 for w in whiches)     # seeds 0/1 hit both outcomes
        assert all(w in ((-1, 0) if mode == 0 else (-1, 1) if mode == 1 else (-1, 0, 1)) for w in whiches)
    with pytest.raises(ValueError):
        rb.perturb(batches[0][0], batches[0][1], 3, 0.1, 0.0)


def test_flop_summary_matches_end_weight_tables():
    from dynmm_b200.fusion import robustness as rb
    d, t, saved = rb.flop_summary([2, 0, 0, 0, 2], fo.DEPTH_ENC_FLOP_R34, fo.TOTAL_FLOP_R34)
    assert abs(d - 0.5 * (fo.DEPTH_ENC_FLOP_R34[0] + fo.DEPTH_ENC_FLOP_R34[4])) < 1e-6
    assert abs(t - 0.5 * (fo.TOTAL_FLOP_R34[0] + fo.TOTAL_FLOP_R34[4])) < 1e-6
    assert abs(saved - 100 * (1 - t / fo.TOTAL_FLOP_R34[4])) < 1e-9
    assert rb.flop_summary([0] * 5, fo.DEPTH_ENC_FLOP_R34, fo.TOTAL_FLOP_R34) == (None, None, None)


def _small_model():
    from dynmm_b200.fusion import SkipGateESANet
    cfg = fo.FusionConfig(height=64, width=64)
    model = SkipGateESANet(height=64, width=64)
    model.load_state_dict(fo.make_state_dict(cfg, 11, 40.0), strict=True)
    model.eval()
    model.hard_gate = True
    return model


def _batches(n=4, b=2):
    g = torch.Generator().manual_seed(21)
    data = []
    for _ in range(n):
        gain = 0.25 + 1.5 * torch.rand(b, 2, generator=g)
        data.append((torch.randn(b, 3, 64, 64, generator=g) * gain[:, :1].view(-1, 1, 1, 1),
                     torch.randn(b, 1, 64, 64, generator=g) * gain[:, 1:].view(-1, 1, 1, 1)))
    return data


def test_run_point_bookkeeping_and_sharding():
    """CPU tensors run the differentiable PyTorch graph of the module (host logic only; the engine is covered by the
    -m gpu tests).  Without noise the union of two rank shards equals the single-process histogram."""
    from dynmm_b200.fusion import robustness as rb
    model, data = _small_model(), _batches()
    seen = []
    full = rb.run_point(model, lambda r: data, -1, 0.0, num_runs=2, on_batch=lambda r, i, p, w: seen.append((r, i, tuple(p.shape), tuple(w.shape))))
    assert full.images == 2 * 4 * 2 and full.batches == 8 and full.noised_batches == 0
    assert sum(full.histogram) == full.images and full.seconds > 0
    assert seen == [(r, i, (2, 40, 64, 64), (2, 5)) for r in range(2) for i in range(4)]
    parts = [rb.run_point(model, lambda r: data, -1, 0.0, num_runs=2, rank=k, world=2) for k in range(2)]
    assert [a + b for a, b in zip(parts[0].histogram, parts[1].histogram)] == full.histogram
    assert parts[0].images + parts[1].images == full.images
    # the pattern of perturbed batches does not depend on the number of ranks
    noisy = rb.run_point(model, lambda r: data, 1, 1.0, num_runs=3)
    shards = [rb.run_point(model, lambda r: data, 1, 1.0, num_runs=3, rank=k, world=2) for k in range(2)]
    assert noisy.noised_batches == shards[0].noised_batches + shards[1].noised_batches
    rb.reduce_point(noisy, model)                       # no process group: only attaches the FLOP summary
    d = noisy.as_dict()
    assert abs(sum(d["gate_branch_fraction"]) - 1) < 1e-9 and d["flop_saved_pct"] is not None
    assert d["total_gflop_per_image"] <= fo.TOTAL_FLOP_R34[4] + 1e-6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynmm_b200.fusion import robustness as rb
        torch.set_num_threads(2)
        model, data = _small_model(), _batches()
        pt = rb.reduce_point(rb.run_point(model, lambda r: data, -1, 0.0, 1, rank, world), model)
        if rank == 0:
            ret["hist"], ret["images"], ret["batches"] = pt.histogram, pt.images, pt.batches
            ret["saved"] = pt.saved_pct
    finally:
        dist.destroy_process_group()


def test_two_rank_sweep_sums_to_single_process():
    from dynmm_b200.fusion import robustness as rb
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    model, data = _small_model(), _batches()
    single = rb.reduce_point(rb.run_point(model, lambda r: data, -1, 0.0, 1), model)
    assert list(ret["hist"]) == single.histogram and ret["images"] == single.images == 8 and ret["batches"] == 4
    assert abs(ret["saved"] - single.saved_pct) < 1e-9
