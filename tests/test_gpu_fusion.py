"""Module-level parity (GPU): the drop-in ``SkipGateESANet`` running on the CUDA
engine against (a) the reference's own outputs stored in tests/golden and (b)
the fp32 CPU oracle on the same seeded weights and inputs.

Stated bf16 tolerance (activations and weights are bf16, accumulation fp32,
~70 convolutions deep): relative L2 error of the logits <= 2e-2 and per-pixel
arg-max agreement >= 99 %.  Gate path is fp32: logits to 1e-4, hard decisions
bit-exact.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_L2_TOL = 2e-2
ARGMAX_AGREE = 0.99


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    yield


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def _build(cfg, seed, gate_scale=40.0):
    from dynmm_b200.fusion import SkipGateESANet
    from oracle import fusion_oracle as fo
    sd = fo.make_state_dict(cfg, seed, gate_scale)
    model = SkipGateESANet(height=cfg.height, width=cfg.width, encoder_rgb=cfg.encoder, encoder_depth=cfg.encoder,
                           encoder_block=cfg.encoder_block, channels_decoder=list(cfg.channels_decoder),
                           nr_decoder_blocks=list(cfg.nr_decoder_blocks),
                           fuse_depth_in_rgb_encoder=cfg.fuse_depth_in_rgb_encoder)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd


def test_engine_matches_reference_golden_vectors(golden_dir):
    """The vectors were produced by the reference's SkipGateESANet itself."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=64, width=96)
    gold = np.load(os.path.join(golden_dir, "fusion_r34_nbt1d_add_64x96.npz"))
    model, sd = _build(cfg, 0, float(gold["gate_scale"]))
    rgb, depth = sample_inputs(1, 4, 64, 96)
    rgb, depth = rgb.cuda(), depth.cuda()

    def check(out, prefix):
        ref = torch.from_numpy(gold[prefix + "_sample"])
        got = out[:, :, ::4, ::4].cpu()
        err = _rel_l2(got, ref)
        assert err <= REL_L2_TOL, f"{prefix}: relative L2 error {err:.4f}"
        agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
        assert agree >= ARGMAX_AGREE, f"{prefix}: arg-max agreement {agree:.4f}"
        return err

    with torch.no_grad():
        for tag, temp, hard in (("soft_t1", 1.0, False), ("hard_t1", 1.0, True), ("soft_t01", 0.1, False)):
            model.temp, model.hard_gate = temp, hard
            out, w = model(rgb, depth, True, True)
            gw = gold[tag + "_weight"]
            if hard:
                np.testing.assert_array_equal(w.cpu().numpy(), gw)          # bit-exact hard decisions
            else:
                np.testing.assert_allclose(w.cpu().numpy(), gw, rtol=2e-3, atol=1e-5)
            check(out, tag + "_out")
        model.baseline = True
        out = model(rgb, depth, True)
        check(out, "baseline_out")
        model.baseline = False
        model.ini_stage = True
        torch.manual_seed(1234)
        out, w = model(rgb, depth, True, True)
        np.testing.assert_array_equal(w.cpu().numpy(), gold["ini_weight"])
        check(out, "ini_out")
        model.ini_stage = False
        eng = model.engine()
        for k in range(5):
            wk = torch.eye(5)[torch.full((4,), k)].cuda()
            out, _ = eng.forward(rgb, depth, weight=wk)
            check(out, f"branch{k}_out")
        # eval-mode call convention without test=True: (out, loss)  (train.py:306 / :316-322)
        model.hard_gate = True
        out, loss = model(rgb, depth)
        ref_loss = (torch.from_numpy(gold["hard_t1_weight"]).mean(0) * torch.tensor(fo.DEPTH_ENC_FLOP_R34)).mean()
        assert abs(loss.item() - ref_loss.item()) < 1e-6


@pytest.mark.parametrize("name", ["fusion_r34_nbt1d_seadd_64x64", "fusion_r50_seadd_decr_64x64",
                                  "fusion_r18_basic_add_64x64"])
def test_engine_matches_reference_golden_other_configs(name, golden_dir):
    """SE-add fusion (the CLI default, args.py:159), ResNet-50 Bottleneck encoders with 'decreasing' decoder
    channels (args.py:105-113,151) and ResNet-18 BasicBlock -- against the reference's own outputs."""
    from tests.test_oracle_golden import CASES
    from oracle.make_golden import sample_inputs
    cfg, seed, b = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    model, sd = _build(cfg, seed, float(gold["gate_scale"]))
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    rgb, depth = rgb.cuda(), depth.cuda()
    with torch.no_grad():
        for tag, temp, hard in (("hard_t1", 1.0, True), ("soft_t1", 1.0, False)):
            model.temp, model.hard_gate = temp, hard
            out, w = model(rgb, depth, True, True)
            if hard:
                np.testing.assert_array_equal(w.cpu().numpy(), gold[tag + "_weight"])
            ref = torch.from_numpy(gold[tag + "_out_sample"])
            err = _rel_l2(out[:, :, ::4, ::4].cpu(), ref)
            assert err <= REL_L2_TOL, f"{name}/{tag}: relative L2 error {err:.4f}"
        eng = model.engine()
        for k in (0, 2, 4):
            wk = torch.eye(5)[torch.full((b,), k)].cuda()
            out, _ = eng.forward(rgb, depth, weight=wk)
            ref = torch.from_numpy(gold[f"branch{k}_out_sample"])
            err = _rel_l2(out[:, :, ::4, ::4].cpu(), ref)
            assert err <= REL_L2_TOL, f"{name}/branch{k}: relative L2 error {err:.4f}"


@pytest.mark.parametrize("variant", ["r18_basic", "r34_nbt1d_odd"])
def test_engine_matches_oracle_other_configs(variant):
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    if variant == "r18_basic":
        cfg, seed, b = fo.FusionConfig(height=64, width=64, encoder="resnet18", encoder_block="BasicBlock"), 5, 2
    else:
        cfg, seed, b = fo.FusionConfig(height=96, width=160), 7, 3
    model, sd = _build(cfg, seed)
    rgb, depth = sample_inputs(seed + 1, b, cfg.height, cfg.width)
    with torch.no_grad():
        ref = fo.forward(sd, cfg, rgb, depth, hard_gate=True)
        model.hard_gate = True
        out, w = model(rgb.cuda(), depth.cuda(), True, True)
    assert torch.equal(w.cpu(), ref["weight"])
    err = _rel_l2(out.cpu(), ref["out"])
    assert err <= REL_L2_TOL, f"relative L2 error {err:.4f}"


def test_full_size_batch_parity_and_skip_equivalence():
    """BASELINE config C2 shape (480x640): engine vs fp32 oracle on 2 images, then
    size-independent properties on the full batch of 8:
      * per-sample independence: sample i of the batch == the same image run alone
        (exact: skipping / slot permutation / multi-sample tiles change no arithmetic);
      * a forced one-hot branch k gives exactly the logits of running branch 4 with
        the skipped stages' gates zeroed -> gated-off depth stages contribute nothing;
      * CUDA-graph replay == eager launches (exact)."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig()
    model, sd = _build(cfg, 0)
    rgb, depth = sample_inputs(21, 8, 480, 640)
    rgb_c, depth_c = rgb.cuda(), depth.cuda()
    eng = model.engine()
    branches = torch.tensor([0, 1, 2, 3, 4, 0, 4, 2])
    wk = torch.eye(5)[branches].cuda()
    with torch.no_grad():
        out8, _ = eng.forward(rgb_c, depth_c, weight=wk)
        ref = fo.forward(sd, cfg, rgb[:2], depth[:2], weight=torch.eye(5)[branches[:2]])
        err = _rel_l2(out8[:2].cpu(), ref["out"])
        assert err <= REL_L2_TOL, f"full-size relative L2 error {err:.4f}"
        agree = (out8[:2].cpu().argmax(1) == ref["out"].argmax(1)).float().mean().item()
        assert agree >= ARGMAX_AGREE, f"arg-max agreement {agree:.4f}"
        for i in (0, 3, 6):
            alone, _ = eng.forward(rgb_c[i:i + 1], depth_c[i:i + 1], weight=wk[i:i + 1])
            assert torch.equal(alone[0], out8[i]), f"sample {i} depends on its batch neighbours"
        # learned hard gate, graph replay vs eager
        model.hard_gate = True
        eager, w_eager = model(rgb_c, depth_c, True, True)
        eager = eager.clone()
        model.use_cuda_graph = True
        g1, w1 = model(rgb_c, depth_c, True, True)
        assert torch.equal(g1, eager) and torch.equal(w1, w_eager)
        perm = torch.arange(7, -1, -1).cuda()
        g2, w2 = model(rgb_c[perm].contiguous(), depth_c[perm].contiguous(), True, True)
        assert torch.equal(w2, w_eager[perm]), "one captured graph must serve every gate outcome"
        assert torch.equal(g2, eager[perm])
        model.use_cuda_graph = False
        ref_w = fo.forward(sd, cfg, rgb[:1], depth[:1], hard_gate=True)["weight"]
        assert torch.equal(w_eager[:1].cpu(), ref_w)


def test_weight_statistics_api():
    """start_weight / end_weight (model_skip_mod_globalgate.py:230-253) without per-forward syncs."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=64, width=96)
    model, _ = _build(cfg, 0)
    rgb, depth = sample_inputs(1, 4, 64, 96)
    model.hard_gate = True
    model.start_weight()
    with torch.no_grad():
        for _ in range(3):
            model(rgb.cuda(), depth.cuda(), True)
    stats = model.end_weight(print_flop=True)
    assert stats is not None and stats[0].sum() == 12
    assert model.weight_list.numel() == 0
    # predict_labels == argmax of forward (fused into the last kernel), and the host pipeline agrees
    from dynmm_b200.fusion import EvalPipeline
    with torch.no_grad():
        ref_labels = model(rgb.cuda(), depth.cuda(), True).argmax(1).to(torch.uint8)
        assert torch.equal(model.predict_labels(rgb.cuda(), depth.cuda()), ref_labels)
        pipe = EvalPipeline(model, 4, 64, 96)
        outs = [l.clone() for l in pipe.run([(rgb.pin_memory(), depth.pin_memory())] * 3)]
    assert len(outs) == 3 and all(torch.equal(o, ref_labels.cpu()) for o in outs)


def test_training_path_uses_custom_gate_ops_and_matches_oracle():
    """Training forward (fp32 autograd graph, custom CUDA DiffSoftmax / gated-blend with custom
    backward) against the CPU oracle in train mode; gate gradients against pure-PyTorch autograd."""
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = fo.FusionConfig(height=64, width=96)
    model, sd = _build(cfg, 0)
    rgb, depth = sample_inputs(1, 4, 64, 96)
    model.train()
    model.temp, model.hard_gate = 1.0, True
    outs, loss = model(rgb.cuda(), depth.cuda())
    with torch.no_grad():
        ref = fo.forward(sd, cfg, rgb, depth, temp=1.0, hard_gate=True, training=True)
    assert len(outs) == 4
    for o, r in zip(outs, ref["out"]):
        assert _rel_l2(o.detach().cpu(), r) < 2e-3
    assert abs(loss.item() - ref["loss"].item()) < 1e-5
    (outs[0].float().mean() + loss).backward()
    g_custom = model.gate_layer.fc.weight.grad.clone()
    assert torch.isfinite(g_custom).all() and g_custom.abs().sum() > 0
    # same graph with the reference formulation of both gate ops
    import dynmm_b200.fusion.modules as M
    model2, _ = _build(cfg, 0)
    model2.train()
    model2.temp, model2.hard_gate = 1.0, True
    orig_blend, orig_ds = M.gated_blend, M.diff_softmax

    def ref_ds(logits, tau=1.0, hard=False, dim=-1):
        y_soft = (logits / tau).softmax(dim)
        if not hard:
            return y_soft
        idx = y_soft.max(dim, keepdim=True)[1]
        return torch.zeros_like(logits).scatter_(dim, idx, 1.0) - y_soft.detach() + y_soft
    try:
        M.gated_blend = lambda r, d, g: (1 - g).view(-1, 1, 1, 1) * r + g.view(-1, 1, 1, 1) * (r + d)
        M.diff_softmax = ref_ds
        outs2, loss2 = model2(rgb.cuda(), depth.cuda())
        (outs2[0].float().mean() + loss2).backward()
    finally:
        M.gated_blend, M.diff_softmax = orig_blend, orig_ds
    g_ref = model2.gate_layer.fc.weight.grad
    assert torch.allclose(g_custom, g_ref, rtol=2e-3, atol=1e-6 + 2e-3 * g_ref.abs().max().item())


def test_miou_on_device_matches_oracle_and_bf16_engine_miou_tolerance():
    """(1) the device arg-max / confusion matrix / mIoU are bit-exact (integers) resp. 1e-12 (fp64) against the
    numpy restatement of eval.py / confusion_matrix.py on the SAME logits -- the fp32 part of the north-star's
    'mIoU within 1e-3' claim is met exactly.  (2) stated bf16 tolerance: mIoU computed from the bf16 engine's
    logits vs mIoU from the fp32 oracle's logits within 3e-2 relative on this adversarial synthetic case (a
    random-init network has near-degenerate class margins, so ~1 % of arg-maxes flip under bf16 rounding and
    every flip is an error against labels derived from the oracle itself).  Labels: the oracle's arg-max
    with 30 % of the pixels re-drawn at random and 10 % void."""
    from dynmm_b200.fusion.metrics import ConfusionMatrix
    from oracle import fusion_oracle as fo
    from oracle import metrics_oracle as mo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=96, width=160)
    model, sd = _build(cfg, 7)
    rgb, depth = sample_inputs(8, 3, cfg.height, cfg.width)
    with torch.no_grad():
        ref = fo.forward(sd, cfg, rgb, depth, hard_gate=True)["out"]
        model.hard_gate = True
        out = model(rgb.cuda(), depth.cuda(), True)
    g = torch.Generator().manual_seed(0)
    label = ref.argmax(1) + 1
    rnd = torch.rand(label.shape, generator=g)
    label = torch.where(rnd < 0.3, torch.randint(1, 41, label.shape, generator=g), label)
    label = torch.where(rnd > 0.9, torch.zeros_like(label), label).to(torch.uint8)
    cm = ConfusionMatrix(40)
    pred = cm.update_from_logits(out, label.cuda(), want_pred=True)
    cm_ref_same_logits = mo.confusion_from_logits(out.cpu().numpy(), label.numpy(), 40)
    assert np.array_equal(cm.confusion_matrix.cpu().numpy(), cm_ref_same_logits)
    assert np.array_equal(pred.cpu().numpy(), out.cpu().numpy().argmax(1).astype(np.uint8))
    miou_dev, iou_dev = cm.compute_miou()
    assert abs(miou_dev - mo.miou(cm_ref_same_logits)) < 1e-12
    miou_fp32 = mo.miou(mo.confusion_from_logits(ref.numpy(), label.numpy(), 40))
    assert abs(miou_dev - miou_fp32) <= 3e-2 * miou_fp32, (miou_dev, miou_fp32)
    agree = (pred.cpu() == ref.argmax(1).to(torch.uint8)).float().mean().item()
    assert agree >= 0.97, agree
    # accumulation over batches
    cm.update_from_logits(out, label.cuda())
    assert np.array_equal(cm.confusion_matrix.cpu().numpy(), 2 * cm_ref_same_logits)


def test_edge_cases_batch_one_all_skipped_and_bad_sizes():
    """Ragged / degenerate inputs: batch 1; a batch in which NO sample keeps any depth stage (every depth-stage
    kernel gets count = 0 and must produce nothing); input sizes that are not a multiple of 32 are refused."""
    from dynmm_b200 import _lib
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=64, width=96)
    model, sd = _build(cfg, 0)
    eng = model.engine()
    rgb, depth = sample_inputs(1, 3, 64, 96)
    with torch.no_grad():
        w0 = torch.eye(5)[torch.zeros(3, dtype=torch.long)].cuda()            # branch 0 everywhere
        out, _ = eng.forward(rgb.cuda(), depth.cuda(), weight=w0)
        ref = fo.forward(sd, cfg, rgb, depth, weight=torch.eye(5)[torch.zeros(3, dtype=torch.long)])["out"]
        assert _rel_l2(out.cpu(), ref) <= REL_L2_TOL
        # batch 1.  (Bit-exact per-sample independence is asserted at full size, where batch 1 and batch 8 pick
        # the same tiling; on tiny maps the batch decides between multi-sample and halo tiles, which changes the
        # fp32 accumulation order, so only the tolerance applies here.)
        one, _ = eng.forward(rgb[:1].cuda(), depth[:1].cuda(), weight=w0[:1])
        assert _rel_l2(one[0].cpu(), ref[0]) <= REL_L2_TOL
        with pytest.raises(_lib.DynmmError):
            eng.forward(torch.zeros(1, 3, 70, 96).cuda(), torch.zeros(1, 1, 70, 96).cuda())
        with pytest.raises(_lib.DynmmError):
            eng.forward(rgb, depth)                                           # CPU tensors


def test_tile_flags_engine_is_bit_identical_at_full_size():
    """DYNMM_TILE_FLAGS=1 (layers overlap through per-tile completion flags instead of kernel boundaries): the whole
    480x640 batch-8 forward gives the bits of the stream-ordered engine, eager and under CUDA-graph replay (30 replays;
    gate outcomes spread over all branches so that every depth stage runs with a partial sample list)."""
    from dynmm_b200 import ops
    from dynmm_b200.fusion.graph import GraphedForward
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig()
    model, sd = _build(cfg, 0)
    rgb, depth = sample_inputs(21, 8, 480, 640)
    rgb_c, depth_c = rgb.cuda(), depth.cuda()
    eng = model.engine()
    wk = torch.eye(5)[torch.tensor([0, 1, 2, 3, 4, 0, 4, 2])].cuda()
    with torch.no_grad():
        assert eng.flag_pool is None or os.environ.get("DYNMM_TILE_FLAGS") == "1"
        eng.flag_pool = None
        ref, _ = eng.forward(rgb_c, depth_c, weight=wk)
        ref = ref.clone()
        model.hard_gate = True
        ref_learned, w_ref = model(rgb_c, depth_c, True, True)
        ref_learned = ref_learned.clone()
        eng.flag_pool = ops.TileFlagPool(rgb_c.device)
        got, _ = eng.forward(rgb_c, depth_c, weight=wk)
        assert eng.flag_pool.off > 1000, "no flags were allocated"
        assert torch.equal(got, ref), "eager forward with tile flags differs"
        modes = dict(temp=1.0, hard_gate=True, baseline=False, ini_stage=False)
        graphed = GraphedForward(eng, rgb_c, depth_c, modes, False)
        for rep in range(30):
            out, w = graphed(rgb_c, depth_c)
            torch.cuda.synchronize()
            assert torch.equal(w, w_ref)
            assert torch.equal(out, ref_learned), f"graph replay {rep} with tile flags differs"
        eng.flag_pool = None


def test_merged_encoder_launches_are_bit_identical_at_full_size():
    """DYNMM_MERGE=1 (stages 2-4: the same layer of both encoders in one launch): the 480x640 batch-8 forward gives the
    bits of the two-stream engine for forced branches (every depth stage runs with a partial sample list) and for
    the learned hard gate under CUDA-graph replay."""
    from dynmm_b200.fusion.graph import GraphedForward
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig()
    model, sd = _build(cfg, 0)
    rgb, depth = sample_inputs(21, 8, 480, 640)
    rgb_c, depth_c = rgb.cuda(), depth.cuda()
    eng = model.engine()
    modes = dict(temp=1.0, hard_gate=True, baseline=False, ini_stage=False)
    with torch.no_grad():
        outs = {}
        for merge in (False, True):
            eng.use_merge = merge
            forced = []
            for branches in ([0, 1, 2, 3, 4, 0, 4, 2], [4] * 8, [0] * 8, [3, 3, 1, 4, 2, 2, 0, 1]):
                wk = torch.eye(5)[torch.tensor(branches)].cuda()
                forced.append(eng.forward(rgb_c, depth_c, weight=wk)[0].clone())
            launches = eng.launches
            graphed = GraphedForward(eng, rgb_c, depth_c, modes, False)
            learned = []
            for rep in range(5):
                o, wgt = graphed(rgb_c, depth_c)
                learned.append((o.clone(), wgt.clone()))
            outs[merge] = (forced, learned, launches)
        eng.use_merge = False
    assert outs[True][2] < outs[False][2] - 40, "merging should remove at least 40 launches per forward"
    for a, b in zip(outs[False][0], outs[True][0]):
        assert torch.equal(a, b), "merged launches changed the logits of a forced-branch forward"
    for (a, wa), (b, wb) in zip(outs[False][1], outs[True][1]):
        assert torch.equal(wa, wb) and torch.equal(a, b), "merged launches changed a graph-replayed forward"
