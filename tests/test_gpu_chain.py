"""dynmm_conv_chain_fwd (a run of NonBottleneck1D convolutions, resnet.py:124-147, as one kernel) against the
per-layer launches of dynmm_conv_igemm_fwd: same accumulation order and rounding, so the comparison is bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layers(c, n_blocks, seed, dev):
    from dynmm_b200 import ops
    g = torch.Generator().manual_seed(seed)
    blocks = []
    for _ in range(n_blocks):
        blk = []
        for i in range(4):
            shape = (c, c, 3, 1) if i % 2 == 0 else (c, c, 1, 3)
            w = torch.randn(shape, generator=g) * (1.5 / (3 * c) ** 0.5)
            blk.append((ops.pack_conv_weight(w.to(dev)), (torch.randn(c, generator=g) * 0.1).to(dev), True))
        blocks.append(blk)
    return blocks


def _reference(x, layers, count=None):
    """Layer by layer on dynmm_conv_igemm_fwd, following the ChainImage layer tuples."""
    from dynmm_b200 import ops
    c = x.shape[3]
    cur, out, last = x, None, None
    for (w, shift, taps_h, relu, residual, store) in layers:
        kh, kw = (3, 1) if taps_h else (1, 3)
        res = {0: None, 1: x, 2: out}[residual]
        cur = ops.conv(cur, w, c_out=c, kh=kh, kw=kw, pad=(kh // 2, kw // 2), shift=shift, relu=bool(relu), residual=res,
                       count=count)
        if store == 1:
            out = cur
        elif store == 2:
            last = cur
    return out, last


@pytest.mark.parametrize("c,h,w,n,n_blocks,drop", [
    (128, 30, 40, 8, 3, False),      # decoder module 2 at 480x640: 10 strips of 3 rows
    (256, 30, 40, 4, 2, True),       # encoder stage 3, last convolution left to the caller
    (128, 60, 80, 8, 1, False),      # three M tiles per strip
    (128, 60, 80, 12, 2, False),     # four M tiles per strip (all of TMEM)
    (128, 15, 20, 8, 3, False),      # strips of 6, 6, 3 rows
    (256, 6, 8, 3, 1, False),        # one strip per sample, no exchange
    (128, 5, 7, 2, 1, True),
    (128, 30, 40, 20, 1, False),     # more strips than SMs: later samples start when earlier ones finish
])
def test_chain_matches_per_layer_launches(c, h, w, n, n_blocks, drop):
    from dynmm_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(c + h)
    x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    layers = ops.nbt1d_chain_layers(_layers(c, n_blocks, 1, dev), drop_last=drop)
    img = ops.ChainImage(layers, c, dev)
    (out, last), = ops.conv_chain([dict(x=x, image=img)])
    ref_out, ref_last = _reference(x, layers)
    torch.cuda.synchronize()
    for got, ref, name in ((out, ref_out, "out"), (last, ref_last, "out_last")):
        assert (got is None) == (ref is None), name
        if ref is None:
            continue
        if h * w >= 128:
            assert torch.equal(got, ref), f"{name}: {(got.float() - ref.float()).abs().max().item()}"
        else:
            # maps smaller than one pixel tile: the per-layer kernel packs several samples into a tile and loops
            # (tap, k chunk) instead of (k chunk, tap) -- same products, another fp32 summation order
            err = (got.float() - ref.float()).norm() / ref.float().norm()
            assert err < 2e-3, f"{name}: relative L2 {err.item()}"


def test_chain_two_jobs_with_count_and_flag_reuse():
    """RGB job (last convolution dropped) + depth job behind a device-side sample count, launched twice on the same
    flags (the kernel leaves them zero) and replayed from a CUDA graph."""
    from dynmm_b200 import ops
    dev = torch.device("cuda")
    c, h, w = 256, 30, 40
    torch.manual_seed(3)
    xr = torch.randn(8, h, w, c, device=dev).to(torch.bfloat16)
    xd = torch.randn(8, h, w, c, device=dev).to(torch.bfloat16)
    lr = ops.nbt1d_chain_layers(_layers(c, 2, 5, dev), drop_last=True)
    ld = ops.nbt1d_chain_layers(_layers(c, 2, 6, dev))
    ir, idp = ops.ChainImage(lr, c, dev), ops.ChainImage(ld, c, dev)
    count = torch.tensor([4], dtype=torch.int32, device=dev)
    units, _ = ops.chain_plan(h, w, c, 16)
    flags = torch.zeros(units + 16, dtype=torch.int32, device=dev)
    jobs = [dict(x=xr, image=ir), dict(x=xd, image=idp, count=count, count_settled=True)]
    ref_r = _reference(xr, lr)
    ref_d = _reference(xd, ld, count=count)
    for _ in range(2):
        (out_r, last_r), (out_d, _) = ops.conv_chain(jobs, flags=flags)
        torch.cuda.synchronize()
        assert int(flags.abs().sum()) == 0
        assert torch.equal(out_r, ref_r[0]) and torch.equal(last_r, ref_r[1])
        assert torch.equal(out_d[:4], ref_d[0][:4])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        (out_r, last_r), (out_d, _) = ops.conv_chain(jobs, flags=flags)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_r, ref_r[0]) and torch.equal(last_r, ref_r[1]) and torch.equal(out_d[:4], ref_d[0][:4])
