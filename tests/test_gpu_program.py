"""Convolution programs (dynmm_conv_program_*): many dependent convolutions in one persistent cooperative
launch must be BIT-IDENTICAL to the same convolutions launched one by one (same tiles, same MMA order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    yield


def _layer(g, cin, cout, kh, kw):
    from dynmm_b200 import ops
    w = torch.randn(cout, cin, kh, kw, device="cuda", generator=g) * (1.0 / (cin * kh * kw) ** 0.5)
    return ops.pack_conv_weight(w), torch.randn(cout, device="cuda", generator=g) * 0.1


def _run_block(x, layers, n, h, w, c_in, c_out, stride, *, count=None, in_map=None, program=False, gated=None,
               gate=None, slot=None):
    """NonBottleneck1D-shaped block with a strided first conv + 1x1 down-sampling residual (resnet.py:124-147)."""
    from dynmm_b200 import ops
    (w1, b1), (w2, b2), (w3, b3), (w4, b4), (wd, bd) = layers
    ho, wo = h // stride, w // stride
    z = lambda hh, ww, c: torch.zeros(n, hh, ww, c, dtype=torch.bfloat16, device="cuda")
    outs = [z(ho, w, c_out), z(ho, w, c_out), z(ho, wo, c_out), z(ho, wo, c_out), z(ho, wo, c_out)]
    kw = dict(count=count, n_out=n)

    def body(step):
        y1 = ops.conv(x, w1, c_out=c_out, kh=3, kw=1, stride=(stride, 1), pad=(1, 0), shift=b1, relu=True, out=outs[0],
                      in_map=in_map, **kw)
        idn = ops.conv(x, wd, c_out=c_out, kh=1, kw=1, stride=(stride, stride), pad=(0, 0), shift=bd, out=outs[4],
                       in_map=in_map, **kw)
        step()
        y2 = ops.conv(y1, w2, c_out=c_out, kh=1, kw=3, stride=(1, stride), pad=(0, 1), shift=b2, relu=True, out=outs[2], **kw)
        step()
        y3 = ops.conv(y2, w3, c_out=c_out, kh=3, kw=1, pad=(1, 0), shift=b3, relu=True, out=outs[3], **kw)
        step()
        y4 = ops.conv(y3, w4, c_out=c_out, kh=1, kw=3, pad=(0, 1), shift=b4, relu=True, residual=idn,
                      out=torch.zeros_like(outs[3]), gated=gated, gate=gate, gated_slot=slot, **kw)
        return y4

    if program:
        with ops.ConvProgram() as prog:
            y = body(prog.next_phase)
        assert prog.n_phases == 4
        return y
    return body(lambda: None)


@pytest.mark.parametrize("shape", [(3, 24, 40, 64, 128, 2), (2, 16, 24, 128, 256, 2), (5, 30, 40, 64, 64, 1),
                                   (8, 120, 160, 64, 64, 1), (8, 30, 40, 256, 512, 2)])
def test_block_program_bit_identical(shape):
    n, h, w, ci, co, stride = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(torch.bfloat16)
    layers = [_layer(g, ci, co, 3, 1), _layer(g, co, co, 1, 3), _layer(g, co, co, 3, 1), _layer(g, co, co, 1, 3),
              _layer(g, ci, co, 1, 1)]
    ref = _run_block(x, layers, n, h, w, ci, co, stride)
    got = _run_block(x, layers, n, h, w, ci, co, stride, program=True)
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
    assert ref.float().abs().max().item() > 0


def test_two_streams_counts_and_gated_add():
    """Depth-like chain (slot order, prefix count, in_map gather) one phase ahead of an RGB-like chain whose last
    conv adds g * depth (model_skip_mod_globalgate.py:279-283); samples beyond `count` are never written."""
    from dynmm_b200 import ops
    n, h, w, c = 6, 24, 32, 64
    g = torch.Generator(device="cuda").manual_seed(2)
    rgb = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    dep = torch.randn(n, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    lw = [_layer(g, c, c, *k) for k in ((3, 1), (1, 3), (3, 1), (1, 3))]
    for active in (0, 1, 4, 6):
        count = torch.tensor([active], dtype=torch.int32, device="cuda")
        perm = torch.tensor([3, 0, 5, 1, 2, 4], dtype=torch.int32, device="cuda")       # slot -> sample
        slot = torch.empty(n, dtype=torch.int32, device="cuda")
        slot[perm.long()] = torch.arange(n, dtype=torch.int32, device="cuda")           # sample -> slot
        gate = torch.zeros(n, device="cuda")
        gate[perm[:active].long()] = 1.0

        def run(program):
            zero = lambda: torch.zeros(n, h, w, c, dtype=torch.bfloat16, device="cuda")
            step = [lambda: None]

            def dconv(x, i, **kw):
                return ops.conv(x, lw[i][0], c_out=c, kh=3 if i % 2 == 0 else 1, kw=1 if i % 2 == 0 else 3,
                                pad=(1, 0) if i % 2 == 0 else (0, 1), shift=lw[i][1], relu=True, count=count, n_out=n,
                                out=zero(), **kw)

            def rconv(x, i, **kw):
                return ops.conv(x, lw[3 - i][0], c_out=c, kh=3 if (3 - i) % 2 == 0 else 1, kw=1 if (3 - i) % 2 == 0 else 3,
                                pad=(1, 0) if (3 - i) % 2 == 0 else (0, 1), shift=lw[3 - i][1], relu=True, out=zero(), **kw)

            def body():
                d = dconv(dep, 0, in_map=perm)
                step[0]()
                d = dconv(d, 1)
                r = rconv(rgb, 0)
                step[0]()
                r = rconv(r, 1, gated=d, gate=gate, gated_slot=slot)
                d2 = dconv(d, 2)
                return r, d, d2

            if program:
                with ops.ConvProgram() as prog:
                    step[0] = prog.next_phase
                    res = body()
                assert prog.n_phases == 3
                return res
            return body()

        ref = run(False)
        got = run(True)
        torch.cuda.synchronize()
        for a, b_ in zip(ref, got):
            assert torch.equal(a, b_), f"active={active}"
        assert torch.equal(got[1][active:], torch.zeros_like(got[1][active:]))        # skipped slots untouched


def test_engine_programs_match_per_launch_path(monkeypatch):
    """Whole gated forward: programs vs one launch per convolution, same weights and inputs -> identical logits."""
    from dynmm_b200.fusion import SkipGateESANet
    from oracle import fusion_oracle as fo
    from oracle.make_golden import sample_inputs
    cfg = fo.FusionConfig(height=96, width=128)
    sd = fo.make_state_dict(cfg, 0, 40.0)
    outs = []
    # at 96x128 the deep stages are smaller than one pixel tile: the chain kernel and the per-layer kernel then sum the
    # same products in different orders (see tests/test_gpu_chain.py), so compare the per-layer forms
    monkeypatch.setenv("DYNMM_CHAIN", "0")
    for flag in ("1", "0"):
        monkeypatch.setenv("DYNMM_PROGRAM", flag)
        model = SkipGateESANet(height=96, width=128)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        model.hard_gate = True
        rgb, depth = sample_inputs(3, 5, 96, 128)
        with torch.no_grad():
            out, wgt = model(rgb.cuda(), depth.cuda(), True, True)
        eng = model.engine()
        assert eng.use_programs == (flag == "1")
        outs.append((out.clone(), wgt.clone(), eng.launches))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][2] < outs[1][2] // 4          # far fewer launches
