"""Fused conv3x1 -> ReLU -> conv1x3 (+BN shift, +residual, ReLU) pair (dynmm_conv_pair_fwd) must be bit-identical to
the two tensor-core convolutions it replaces (same taps, accumulation order and bf16 rounding of the intermediate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    yield


def _weights(g, kh, kw):
    from dynmm_b200 import ops
    w = torch.randn(64, 64, kh, kw, device="cuda", generator=g) * (1.0 / (64 * kh * kw) ** 0.5)
    return ops.pack_conv_weight(w), torch.randn(64, device="cuda", generator=g) * 0.1


@pytest.mark.parametrize("shape", [(2, 16, 28), (3, 24, 40), (1, 9, 15), (2, 120, 160), (5, 30, 41)])
@pytest.mark.parametrize("residual", [False, True])
def test_pair_bit_identical(shape, residual):
    from dynmm_b200 import ops
    n, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, h, w, 64, device="cuda", generator=g).to(torch.bfloat16)
    (w1, b1), (w2, b2) = _weights(g, 3, 1), _weights(g, 1, 3)
    res = torch.randn(n, h, w, 64, device="cuda", generator=g).to(torch.bfloat16) if residual else None
    y = ops.conv(x, w1, c_out=64, kh=3, kw=1, pad=(1, 0), shift=b1, relu=True)
    ref = ops.conv(y, w2, c_out=64, kh=1, kw=3, pad=(0, 1), shift=b2, relu=True, residual=res)
    got = ops.conv_pair(x, w1, b1, w2, b2, residual=res, relu2=True)
    torch.cuda.synchronize()
    assert ref.float().abs().max().item() > 0
    assert torch.equal(ref, got), (ref.float() - got.float()).abs().max().item()


def test_pair_count_and_maps():
    """Depth-encoder form: slot order with a prefix count, gathered input and residual samples."""
    from dynmm_b200 import ops
    n, h, w = 6, 16, 28
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(n, h, w, 64, device="cuda", generator=g).to(torch.bfloat16)
    (w1, b1), (w2, b2) = _weights(g, 3, 1), _weights(g, 1, 3)
    perm = torch.tensor([3, 0, 5, 1, 2, 4], dtype=torch.int32, device="cuda")
    for active in (0, 1, 4, 6):
        count = torch.tensor([active], dtype=torch.int32, device="cuda")
        y = ops.conv(x, w1, c_out=64, kh=3, kw=1, pad=(1, 0), shift=b1, relu=True, count=count, in_map=perm, n_out=n,
                     out=torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device="cuda"))
        ref = ops.conv(y, w2, c_out=64, kh=1, kw=3, pad=(0, 1), shift=b2, relu=True, residual=x, res_map=perm,
                       count=count, n_out=n, out=torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device="cuda"))
        got = ops.conv_pair(x, w1, b1, w2, b2, residual=x, relu2=True, count=count, in_map=perm, res_map=perm,
                            n_out=n, out=torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device="cuda"))
        torch.cuda.synchronize()
        assert torch.equal(ref, got), f"active={active}"


def test_pair_rotated_store_order_is_bit_identical():
    """The conflict-avoiding chunk order in epilogue 1 is the default since round 2 (DYNMM_PAIR_ROT=0 selects the straight
    order): the switch is read once per process, so the bit-identity tests above are re-run in a child process with the
    OTHER setting -- both orders must give the bits of two separate convolutions."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, DYNMM_PAIR_ROT="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_pair.py"), "-q", "-x",
                        "-m", "gpu", "-k", "bit_identical and not rotated or count_and_maps", "-p", "no:cacheprovider"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout
