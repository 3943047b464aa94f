"""CPU: the torch mirrors of the reference's three MLP modules (gomavatar_b200/modules.py) load the reference modules'
state dicts unchanged and reproduce their outputs (tests/golden/golden_modules.npz, produced by running the reference's
own modules: oracle/make_golden.py::modules_golden)."""
import os

import numpy as np
import pytest
import torch

from gomavatar_b200.modules import NonRigidModule, PoseRefinementModule, ShadowModule, hann_window, posenc

t = torch.from_numpy


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_modules.npz"))


def _cfg(gold, name):
    pre = f"cfg.{name}."
    cfg = {k[len(pre):]: (int(v) if float(v).is_integer() else float(v)) for k, v in gold.items() if k.startswith(pre)}
    cfg["skips"] = [4]
    return cfg


def _load(mod, gold, prefix):
    sd = {k[len(prefix) + 1:]: t(v) for k, v in gold.items() if k.startswith(prefix + ".")}
    missing, unexpected = mod.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_shadow_module_matches_reference(gold):
    m = ShadowModule(_cfg(gold, "shadow_module"))
    _load(m, gold, "shadow")
    with torch.no_grad():
        out = m(t(gold["shadow_in"]))
    np.testing.assert_allclose(out.numpy(), gold["shadow_out"], rtol=1e-5, atol=1e-6)


def test_non_rigid_module_matches_reference_across_the_hann_window(gold):
    m = NonRigidModule(_cfg(gold, "non_rigid"))
    _load(m, gold, "non_rigid")
    xyz, pv = t(gold["non_rigid_xyz"]), t(gold["non_rigid_posevec"])
    for it in (150000, 163000, 187500, 300000):
        with torch.no_grad():
            out, R, S = m(xyz, pv, it, R=None, S=None)
        assert R is None and S is None
        np.testing.assert_allclose(out.numpy(), gold[f"non_rigid_out_{it}"], rtol=1e-5, atol=2e-6)
    w = hann_window(6, 150000, 150000, 200000)
    assert float(w.abs().sum()) == 0.0                               # nothing passes at kick-in
    assert torch.allclose(hann_window(6, 300000, 150000, 200000), torch.ones(6))


def test_pose_refinement_module_matches_reference(gold):
    m = PoseRefinementModule(_cfg(gold, "pose_refinement"))
    _load(m, gold, "pose_refinement")
    with torch.no_grad():
        out = m(t(gold["pose_in"]))
    assert out.shape == (2, 24, 3, 3)
    np.testing.assert_allclose(out.numpy(), gold["pose_out"], rtol=1e-5, atol=1e-6)
    assert torch.equal(out[:, 0], torch.eye(3).expand(2, 3, 3))       # the root joint is never refined


def test_initialisation_follows_the_reference_scheme():
    torch.manual_seed(0)
    for m in (ShadowModule({"multires": 6, "mlp_width": 128, "mlp_depth": 3, "skips": [4]}),
              NonRigidModule({"multires": 6, "mlp_width": 128, "mlp_depth": 6, "skips": [4], "condition_code_size": 69})):
        last = m.block_mlps[-1]
        assert float(last.weight.abs().max()) <= 1e-5 and float(last.bias.abs().max()) == 0.0
        first = m.block_mlps[0]
        bound = np.sqrt(2.0) * np.sqrt(6.0 / (first.in_features + first.out_features))     # Xavier-uniform, ReLU gain
        assert float(first.weight.abs().max()) <= bound + 1e-6 and float(first.weight.abs().max()) > 0.8 * bound
    assert posenc(torch.zeros(2, 3), 6).shape == (2, 39)
