"""ModalityDynMM on the host: module surface vs the CPU oracle (same state_dict), and the
reference's FLOP constants as architecture pins."""
import pytest
import torch

from oracle import modality_oracle as mo


def test_flop_constants_pin_the_architectures():
    e1, e2 = mo.imdb_mmacs()
    assert abs(e1 - 1.25261) < 1e-5 and abs(e2 - 10.86908) < 1e-5          # imdb_dyn.py:66
    e1, e2 = mo.mosei_mmacs(50)
    assert abs(e1 - 135.13226) < 1e-5 and abs(e2 - 320.03205) < 1e-5       # affect_dyn.py:126


def _randomize_bn(model, g):
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


@pytest.mark.parametrize("hard", [False, True])
def test_imdb_module_matches_oracle(hard):
    from dynmm_b200.modality import DynMMNet
    torch.manual_seed(0)
    model = DynMMNet(pretrain=False, freeze=True).eval()
    g = torch.Generator().manual_seed(1)
    _randomize_bn(model, g)
    model.hard_gate = hard
    inputs = [torch.randn(128, 300, generator=g), torch.randn(128, 4096, generator=g)]
    with torch.no_grad():
        out, reg = model(inputs)
        ref_out, ref_reg, ref_w = mo.imdb_forward(model.state_dict(), inputs, 1.0, hard)
    assert out.shape == (128, 23)
    torch.testing.assert_close(out, ref_out, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(reg, ref_reg, rtol=1e-5, atol=1e-6)
    model.infer_mode = 2
    with torch.no_grad():
        p, zero = model(inputs)
        ref_p, _ = mo.imdb_forward(model.state_dict(), inputs, 1.0, hard, infer_mode=2)
    assert zero == 0
    torch.testing.assert_close(p, ref_p, rtol=1e-4, atol=1e-5)
    # freeze=True leaves only the gate trainable (imdb_dyn.py:52-57,68-70)
    assert {n.split(".")[0] for n, p_ in model.named_parameters() if p_.requires_grad} == {"gate"}


def test_mosei_module_matches_oracle():
    """BASELINE config C1: soft gate, 2 experts, batch 32, T=50, CPU."""
    from dynmm_b200.modality import DynMMNetV2
    torch.manual_seed(0)
    model = DynMMNetV2(temp=1.0, hard_gate=False, freeze=True, model_name_list=None).eval()
    g = torch.Generator().manual_seed(2)
    feats = [torch.randn(32, 50, d, generator=g) for d in (35, 74, 300)]
    lens = [torch.full((32,), 50)] * 3
    with torch.no_grad():
        out, reg = model([feats, lens])
        ref_out, ref_reg, _ = mo.mosei_forward(model.state_dict(), [feats, lens], 1.0, False)
    assert out.shape == (32, 1)
    torch.testing.assert_close(out, ref_out, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(reg, ref_reg, rtol=1e-4, atol=1e-6)
    model.infer_mode = -1
    with torch.no_grad():
        out, _ = model([feats, lens])
        ref_out, _, _ = mo.mosei_forward(model.state_dict(), [feats, lens], 1.0, False, infer_mode=-1)
    torch.testing.assert_close(out, ref_out, rtol=2e-4, atol=2e-5)


def test_gate_gradient_flows_through_straight_through_estimator():
    from dynmm_b200.modality import DynMMNet
    torch.manual_seed(0)
    model = DynMMNet(pretrain=False, freeze=True).train()
    model.hard_gate = True
    inputs = [torch.randn(16, 300), torch.randn(16, 4096)]
    out, reg = model(inputs)
    (out.square().mean() + 0.1 * reg).backward()
    assert model.gate.fc2.weight.grad.abs().sum() > 0
    assert model.text_encoder.fc.weight.grad is None


def test_multibench_import_aliases_and_pickle_roundtrip(tmp_path):
    from dynmm_b200.modality import DynMMNet
    from dynmm_b200.modality.compat import install_aliases
    install_aliases()
    from unimodals.common_models import MLP, MaxOut_MLP           # noqa: F401  (MultiBench import path)
    from fusions.common_fusions import Concat                      # noqa: F401
    from training_structures.Supervised_Learning import MMDL       # noqa: F401
    from src.models.model_skip_mod_globalgate import SkipGateESANet  # noqa: F401  (build_model.py:11-12)
    from src.models.model_skip_mod import SkipESANet               # noqa: F401
    model = DynMMNet(pretrain=False, freeze=False)
    path = tmp_path / "m.pt"
    torch.save(model, path)                                        # Supervised_Learning.py:208 saves whole modules
    again = torch.load(path, weights_only=False)
    assert sorted(again.state_dict()) == sorted(model.state_dict())
