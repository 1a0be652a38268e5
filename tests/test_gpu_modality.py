"""ModalityDynMM on the GPU: gate + mix on the custom kernels, experts really skipped under
hard gates; compared with the CPU oracle (fp32; tolerance 1e-4 relative, TF32 disabled)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    from dynmm_b200 import _lib
    _lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def test_imdb_hard_gate_routes_rows_and_matches_oracle():
    """BASELINE config C4: MM-IMDB late-fusion DynMM, hard gate, batch 128."""
    from dynmm_b200.modality import DynMMNet
    from oracle import modality_oracle as mo
    torch.manual_seed(0)
    model = DynMMNet(pretrain=False, freeze=True).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    inputs = [torch.randn(128, 300, generator=g), torch.randn(128, 4096, generator=g)]
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    model.hard_gate = True
    calls = []
    orig = model.branch3.forward
    model.branch3.forward = lambda x: (calls.append(x[0].shape[0]), orig(x))[1]
    with torch.no_grad():
        out, reg = model([t.cuda() for t in inputs])
        ref_out, ref_reg, ref_w = mo.imdb_forward(sd, inputs, 1.0, True)
    torch.cuda.synchronize()
    k1 = int(ref_w[:, 1].sum().item())
    assert model.last_route_counts == [128 - k1, k1]
    assert calls == ([k1] if k1 else []), "the expensive expert must only see the rows routed to it"
    torch.testing.assert_close(out.cpu(), ref_out, rtol=1e-4, atol=1e-4)
    assert abs(reg.item() - ref_reg.item()) < 1e-6
    # soft gate: both experts on every row, custom mix kernel
    model.hard_gate = False
    with torch.no_grad():
        out, reg = model([t.cuda() for t in inputs])
        ref_out, ref_reg, _ = mo.imdb_forward(sd, inputs, 1.0, False)
    torch.testing.assert_close(out.cpu(), ref_out, rtol=1e-4, atol=1e-4)


def test_mix_autograd_matches_pytorch():
    from dynmm_b200.modality import mix
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.softmax(torch.randn(64, 3, device="cuda", generator=g), 1).requires_grad_(True)
    preds = [torch.randn(64, 7, device="cuda", generator=g).requires_grad_(True) for _ in range(3)]
    up = torch.randn(64, 7, device="cuda", generator=g)
    out = mix(w, preds)
    out.backward(up)
    got = [w.grad.clone()] + [p.grad.clone() for p in preds]
    w.grad = None
    for p in preds:
        p.grad = None
    ref = sum(w[:, e:e + 1] * preds[e] for e in range(3))
    ref.backward(up)
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
    for a, b in zip(got, [w.grad] + [p.grad for p in preds]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


def test_mosei_soft_gate_matches_oracle():
    """BASELINE config C1 shapes (soft gate, 2 experts, batch 32, T=50) on the GPU."""
    from dynmm_b200.modality import DynMMNetV2
    from oracle import modality_oracle as mo
    torch.manual_seed(0)
    model = DynMMNetV2(temp=1.0, hard_gate=False, freeze=True, model_name_list=None).eval()
    g = torch.Generator().manual_seed(2)
    feats = [torch.randn(32, 50, d, generator=g) for d in (35, 74, 300)]
    lens = [torch.full((32,), 50)] * 3
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    with torch.no_grad():
        out, reg = model([[f.cuda() for f in feats], lens])
        ref_out, ref_reg, ref_w = mo.mosei_forward(sd, [feats, lens], 1.0, False)
        torch.testing.assert_close(out.cpu(), ref_out, rtol=1e-3, atol=1e-4)
        model.hard_gate = True
        out, reg = model([[f.cuda() for f in feats], lens])
        ref_out, _, ref_w = mo.mosei_forward(sd, [feats, lens], 1.0, True)
    assert sum(model.last_route_counts) == 32
    torch.testing.assert_close(out.cpu(), ref_out, rtol=1e-3, atol=1e-4)
