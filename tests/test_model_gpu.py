"""GPU parity, end to end: ``gomavatar_b200.model.Model.forward`` against the outputs of the REFERENCE's unmodified
``Model.forward`` + ``Renderer.forward`` (tests/golden/golden_model.npz; rasterizer stubbed by the C oracle when the
golden was made) and gradients against the oracle chain."""
import os

import numpy as np
import pytest
import torch

from gomavatar_b200 import synthetic as S
from oracle import camera as Cam
from oracle import geometry as G
from oracle import raster as R

pytestmark = pytest.mark.gpu
t = torch.from_numpy
DEV = "cuda:0"


def _model_from_golden(g):
    from gomavatar_b200.model import Model, default_model_cfg
    sc = S.make_humanoid(int(g["n_faces"]), seed=int(g["scene_seed"]))
    img = int(g["img_size"])
    m = Model(default_model_cfg(img_size=(img, img)), sc.canonical_info()).to(DEV)
    with torch.no_grad():
        m.so3.copy_(t(g["so3"])); m.scale.copy_(t(g["scale"])); m.appearance_module.appearance.copy_(t(g["appearance"]))
    m.train()
    return m, sc


def _close_image(got, ref, what):
    err = np.abs(got - ref)
    assert (err > 1e-4 * np.abs(ref) + 1e-5).mean() <= 5e-4, what
    assert err.max() < 2e-2, (what, err.max())


def test_forward_matches_reference_model_forward(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    m, sc = _model_from_golden(g)
    c = lambda k, sl=slice(None): t(g[k][sl]).to(DEV)
    # one frame at a time, exactly like train.py:317-323
    for b in range(3):
        s = slice(b, b + 1)
        rgbs, masks, out = m(c("K", s), c("E", s), c("cnl_gtfms", s), c("dst_Rs", s), c("dst_Ts", s),
                             dst_posevec=c("dst_posevec", s), i_iter=0, bgcolor=c("bgcolor", s))
        assert rgbs.shape == (1, 64, 64, 3) and masks.shape == (1, 64, 64)
        _close_image(rgbs[0].detach().cpu().numpy(), g[f"rgbs_{b}"][0], f"rgbs frame {b}")
        _close_image(masks.detach().cpu().numpy(), g[f"masks_{b}"], f"masks frame {b}")
        _close_image(out["albedo"].detach().cpu().numpy(), g[f"albedo_{b}"], f"albedo frame {b}")
    # and all three frames in one batched call
    rgbs, masks, _ = m(c("K"), c("E"), c("cnl_gtfms"), c("dst_Rs"), c("dst_Ts"), dst_posevec=c("dst_posevec"), i_iter=0)
    for b in range(3):
        _close_image(rgbs[b].detach().cpu().numpy(), g[f"rgbs_{b}"][0], f"batched rgbs frame {b}")
    # rigid test-time-pose branch (train_pose.py)
    s = slice(0, 1)
    rgbs, _, _ = m(c("K", s), c("E", s), c("cnl_gtfms", s), c("dst_Rs", s), c("dst_Ts", s), dst_posevec=c("dst_posevec", s),
                   i_iter=0, global_R=t(g["global_R"]).to(DEV), global_T=t(g["global_T"]).to(DEV))
    _close_image(rgbs[0].detach().cpu().numpy(), g["rigid_rgbs"][0], "rigid branch")


def test_camera_kernel_matches_reference_host_math(golden_dir):
    from gomavatar_b200.camera import camera_from_KE
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    view, proj, tanfov, campos = camera_from_KE(t(g["K"]).to(DEV), t(g["E"]).to(DEV), 64, 64, with_campos=True)
    for b in range(3):
        assert np.array_equal(view[b].cpu().numpy(), g[f"pass0_view_{b}"])
        np.testing.assert_allclose(proj[b].cpu().numpy(), g[f"pass0_proj_{b}"], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(tanfov[b].cpu().numpy().astype(np.float64), g[f"pass0_tanfov_{b}"], rtol=1e-7)
        np.testing.assert_allclose(campos[b].cpu().numpy(), g[f"pass0_campos_{b}"], atol=1e-5)


def test_gradients_match_oracle_chain(golden_dir):
    """d(loss)/d(vertices, so3, scale, appearance, dst_Rs, dst_Ts) through LBS -> frame -> cov -> splat, vs the oracle:
    autograd (float32 torch restatement of model.py:212-250) chained with the C rasterizer's explicit backward."""
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    m, sc = _model_from_golden(g)
    b = 1
    H = W = 64
    rng = np.random.default_rng(5)
    dL_rgb = rng.normal(size=(1, H, W, 3)).astype(np.float32)
    dL_mask = rng.normal(size=(1, H, W)).astype(np.float32)
    s = slice(b, b + 1)
    c = lambda k: t(g[k][s]).to(DEV)
    dR, dT = c("dst_Rs").requires_grad_(True), c("dst_Ts").requires_grad_(True)
    rgbs, masks, _ = m(c("K"), c("E"), c("cnl_gtfms"), dR, dT, i_iter=0)
    ((rgbs * t(dL_rgb).to(DEV)).sum() + (masks * t(dL_mask).to(DEV)).sum()).backward()
    # oracle
    ov = t(sc.vertices.T.copy()).requires_grad_(True)
    ow, os_ = t(g["so3"]).requires_grad_(True), t(g["scale"]).requires_grad_(True)
    oR, oT = t(g["dst_Rs"][b]).requires_grad_(True), t(g["dst_Ts"][b]).requires_grad_(True)
    _, xyz, cov = G.pose_geometry(ov, t(sc.faces), t(sc.lbs_weights), ow, os_, t(g["cnl_gtfms"][b]), oR, oT)
    cov6 = G.pack_cov6(cov)
    st = Cam.raster_settings_from_KE(g["K"][b], g["E"][b], (W, H))
    app = g["appearance"].T
    feat = np.concatenate([app, np.ones_like(app[:, :1])], 1)
    fwd = R.forward(xyz.detach().numpy(), cov6.detach().numpy(), feat, np.ones(len(app), np.float32), st.viewmatrix,
                    st.projmatrix, st.tanfovx, st.tanfovy, np.zeros(4, np.float32), H, W)
    dL = np.concatenate([dL_rgb[0].transpose(2, 0, 1), dL_mask], 0)
    gr = R.backward(fwd, dL)
    ((xyz * t(gr["means3D"])).sum() + (cov6 * t(gr["cov6"])).sum()).backward()
    pairs = [("vertices", m.vertices.grad, ov.grad), ("so3", m.so3.grad, ow.grad), ("scale", m.scale.grad, os_.grad),
             ("appearance", m.appearance_module.appearance.grad, t(gr["colors"][:, :3].T.copy())),
             ("dst_Rs", dR.grad[0], oR.grad), ("dst_Ts", dT.grad[0], oT.grad)]
    for name, got, ref in pairs:
        got, ref = got.cpu().numpy(), ref.numpy()
        rel = np.abs(got - ref).max() / np.abs(ref).max()
        assert rel < 1e-3, (name, rel)


def test_state_dict_keys_match_reference_checkpoint_layout(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    m, _ = _model_from_golden(g)
    keys = set(m.state_dict().keys())
    assert {"faces", "lbs_weights", "vertices", "so3", "scale", "appearance_module.appearance",
            "appearance_module.bg_col"} <= keys
    assert m.vertices.shape[0] == 3 and m.so3.shape[0] == 3 and m.lbs_weights.shape[0] == 25
