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


def test_rigid_pose_gradients_match_oracle_chain(golden_dir):
    """d(loss)/d(global_R, global_T) — the only reason those arguments exist: the reference's test-time pose fit optimises
    exactly them (models/model.py:218-221, train_pose.py:247-254, RodriguesModule utils/network_util.py:66-92) — through
    Rodrigues -> rigid transform -> face frame -> covariance -> splat, against the oracle chain (float32 torch restatement
    + the C rasterizer's explicit backward).  Also d/d(vertices) on the same path, which now runs through the rotation."""
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    m, sc = _model_from_golden(g)
    b = 0
    H = W = 64
    rng = np.random.default_rng(9)
    dL_rgb = rng.normal(size=(1, H, W, 3)).astype(np.float32)
    dL_mask = rng.normal(size=(1, H, W)).astype(np.float32)
    s = slice(b, b + 1)
    c = lambda k: t(g[k][s]).to(DEV)
    for gR0, gT0 in ((g["global_R"], g["global_T"]), (np.zeros(3, np.float32), np.zeros(3, np.float32))):   # theta -> sqrt(1e-5): the eps branch
        m.zero_grad(set_to_none=True)
        gR, gT = t(gR0.copy()).to(DEV).requires_grad_(True), t(gT0.copy()).to(DEV).requires_grad_(True)
        rgbs, masks, _ = m(c("K"), c("E"), c("cnl_gtfms"), c("dst_Rs"), c("dst_Ts"), i_iter=0, global_R=gR, global_T=gT)
        ((rgbs * t(dL_rgb).to(DEV)).sum() + (masks * t(dL_mask).to(DEV)).sum()).backward()
        ov = t(sc.vertices.T.copy()).requires_grad_(True)
        oR, oT = t(gR0.copy()).requires_grad_(True), t(gT0.copy()).requires_grad_(True)
        _, xyz, cov = G.pose_geometry(ov, t(sc.faces), t(sc.lbs_weights), t(g["so3"]), t(g["scale"]), t(g["cnl_gtfms"][b]),
                                      t(g["dst_Rs"][b]), t(g["dst_Ts"][b]), global_R=oR, global_T=oT)
        cov6 = G.pack_cov6(cov)
        st = Cam.raster_settings_from_KE(g["K"][b], g["E"][b], (W, H))
        app = g["appearance"].T
        feat = np.concatenate([app, np.ones_like(app[:, :1])], 1)
        fwd = R.forward(xyz.detach().numpy(), cov6.detach().numpy(), feat, np.ones(len(app), np.float32), st.viewmatrix,
                        st.projmatrix, st.tanfovx, st.tanfovy, np.zeros(4, np.float32), H, W)
        gr = R.backward(fwd, np.concatenate([dL_rgb[0].transpose(2, 0, 1), dL_mask], 0))
        ((xyz * t(gr["means3D"])).sum() + (cov6 * t(gr["cov6"])).sum()).backward()
        assert float(oR.grad.abs().max()) > 0 and float(oT.grad.abs().max()) > 0
        for name, got, ref in (("global_R", gR.grad, oR.grad), ("global_T", gT.grad, oT.grad), ("vertices", m.vertices.grad, ov.grad)):
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


@pytest.mark.parametrize("shadow_width", [64, 128])        # 128: the tcgen05 FusedShadowModule; 64: the torch module
def test_full_model_with_all_reference_modules_matches_oracle_composition(shadow_width):
    """Model built from a reference-shaped cfg (pose refinement, non-rigid, mesh normal renderer, shadow MLP all on):
    rgbs = albedo * shading, masks, normal map and soft normal mask against the CPU composition of the oracles, and the
    backward reaches every parameter group (reference models/model.py:184-303)."""
    import copy
    from gomavatar_b200.model import Model
    from oracle import mesh_raster as MR
    W = H = 64
    cfg = {"img_size": [W, H], "eval_mode": False,
           "canonical_geometry": {"sigma": 1e-3, "radius_scale": 1.0, "deform_scale": True, "deform_so3": True},
           "appearance": {"color_init": 0.5},
           "pose_refinement": {"name": "basic", "embedding_size": 69, "total_bones": 24, "mlp_width": 64, "mlp_depth": 2,
                               "refine_root": False, "refine_t": False, "kick_in_iter": 0},
           "non_rigid": {"name": "basic", "condition_code_size": 69, "mlp_width": 64, "mlp_depth": 3, "skips": [4], "multires": 6,
                         "i_embed": 0, "kick_in_iter": 0, "full_band_iter": 10},
           "normal_renderer": {"name": "mesh", "soft_mask": True, "sigma": 1e-5},
           "shadow_module": {"name": "basic", "mlp_width": shadow_width, "mlp_depth": 3, "skips": [4], "multires": 6}}
    sc = S.make_humanoid(2000, seed=0)
    fr = S.make_frames(sc, 1, img_size=(W, H), seed=3)
    pr = S.make_params(sc, seed=1)
    torch.manual_seed(4)
    m = Model(cfg, sc.canonical_info())
    assert m.pose_refinement_module is not None and m.non_rigid_module is not None and m.normal_renderer is not None and m.shadow_module is not None
    with torch.no_grad():
        m.so3.copy_(t(pr["so3"])); m.scale.copy_(t(pr["scale"])); m.appearance_module.appearance.copy_(t(pr["appearance"]))
        m.pose_refinement_module.block_mlps[-1].weight.normal_(0, 2e-3)          # a visible pose correction
        m.non_rigid_module.block_mlps[-1].weight.normal_(0, 2e-4)                # a visible (sub-centimetre) offset
        m.shadow_module.block_mlps[-1].weight.normal_(0, 0.3)
    from gomavatar_b200.modules import ShadowModule
    from gomavatar_b200.shadow import FusedShadowModule
    assert isinstance(m.shadow_module, FusedShadowModule) == (shadow_width == 128)
    cpu = copy.deepcopy(m)                                                       # same weights for the CPU composition
    cpu.shadow_module = ShadowModule(cfg["shadow_module"])                       # (the fused module has no CPU path)
    cpu.shadow_module.load_state_dict(m.shadow_module.state_dict())
    m = m.to(DEV).train()
    d = {k: t(v).to(DEV) for k, v in fr.items()}
    rgbs, masks, out = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"], dst_posevec=d["dst_posevec"], i_iter=7)
    for k in ("colors", "face_connectivity", "target_edge_length", "albedo", "normal", "normal_mask", "shadow"):
        assert k in out, k
    # ---- CPU composition
    with torch.no_grad():
        pv = t(fr["dst_posevec"])
        dR = torch.matmul(t(fr["dst_Rs"]).reshape(-1, 3, 3), cpu.pose_refinement_module(pv).reshape(-1, 3, 3)).reshape(1, 24, 3, 3)
        v_pose, _, _ = cpu.non_rigid_module(cpu.vertices[None], pv, 7)
        v_obs, xyz, cov = G.pose_geometry(v_pose[0], t(sc.faces), t(sc.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                          t(fr["cnl_gtfms"][0]), dR[0], t(fr["dst_Ts"][0]))
        st = Cam.raster_settings_from_KE(fr["K"][0], fr["E"][0], (W, H))
        app = pr["appearance"].T
        feat = np.ascontiguousarray(np.concatenate([app, np.ones_like(app[:, :1])], 1), dtype=np.float32)
        f = R.forward(xyz.numpy(), G.pack_cov6(cov).numpy(), feat, np.ones(len(app), np.float32), st.viewmatrix, st.projmatrix,
                      st.tanfovx, st.tanfovy, np.zeros(4, np.float32), H, W)
        albedo = t(f["color"].transpose(1, 2, 0).copy())
        nm, nmask = MR.render(v_obs.T.contiguous(), t(sc.faces).long(), t(fr["K"][0]), t(fr["E"][0]), H, W, training=True, sigma_cfg=1e-5)
        shade = cpu.shadow_module(nm.reshape(1, H * W, 3)).reshape(H, W, 1) * 2
        ref_rgb = albedo[..., :3] * shade
    got = rgbs[0].detach().cpu().numpy()
    err = np.abs(got - ref_rgb.numpy())
    assert (err > 1e-4 * np.abs(ref_rgb.numpy()) + 2e-5).mean() <= 3e-3 and err.max() < 5e-2, (float(err.max()), float((err > 1e-4).mean()))
    _close_image(masks[0].detach().cpu().numpy(), albedo[..., 3].numpy(), "mask")
    dn = np.abs(out["normal"][0].detach().cpu().numpy() - nm.numpy()).max(-1)
    assert (dn > 1e-4).mean() <= 3e-3
    dm = np.abs(out["normal_mask"][0].detach().cpu().numpy() - nmask.numpy())
    assert (dm > 2e-3).mean() <= 3e-3
    # ---- backward reaches every parameter group
    rng = np.random.default_rng(0)
    loss = (rgbs * t(rng.normal(size=(1, H, W, 3)).astype(np.float32)).to(DEV)).sum() + masks.sum() + out["normal_mask"].sum()
    loss.backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        assert float(p.grad.abs().max()) > 0, name
    names = {g["name"] for g in m.get_param_groups({"lr": {"appearance": 1e-3, "canonical_geometry": 1e-3, "canonical_geometry_xyz": 1e-3,
                                                           "non_rigid": 1e-3, "pose_refinement": 1e-4, "shadow": 1e-3}})}
    assert {"appearance", "canonical_geometry_xyz", "canonical_geometry", "non_rigid", "pose_refinement", "shadow"} <= names


def test_hot_path_is_cuda_graph_capturable(golden_dir):
    """No host sync, no allocation outside torch's allocator, every launch on the caller's stream: forward + photometric
    loss + backward of the hot path captured ONCE in a CUDA graph and replayed on new inputs equals the eager result
    (the reference's rasterizer reads num_rendered back to the host twice per frame and cannot be captured)."""
    from gomavatar_b200.losses import photometric_l1
    g = np.load(os.path.join(golden_dir, "golden_model.npz"))
    m, sc = _model_from_golden(g)
    m.strict_raster = False                                   # overflow is a device flag, checked after the replay
    B, H, W = 2, 64, 64
    keys = ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "bgcolor")
    static = {k: t(g[k][:B]).to(DEV).clone() for k in keys}
    rng = np.random.default_rng(3)
    gt = t(rng.random((B, H, W, 3)).astype(np.float32)).to(DEV)
    gtm = t((rng.random((B, H, W)) > 0.5).astype(np.float32)).to(DEV)
    params = [m.vertices, m.so3, m.scale, m.appearance_module.appearance]
    for p in params:
        p.grad = torch.zeros_like(p)

    def step():
        for p in params:
            p.grad.zero_()
        rgb, mask, _ = m(static["K"], static["E"], static["cnl_gtfms"], static["dst_Rs"], static["dst_Ts"])
        _, l_rgb, l_mask = photometric_l1(rgb, mask, static["bgcolor"], gt, gtm)
        loss = l_rgb + 5.0 * l_mask
        loss.backward()
        return loss.detach(), rgb.detach()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                             # warm-up on a side stream, as torch's capture recipe asks
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_g, rgb_g = step()
    # new inputs: frames 1..2 instead of 0..1
    for k in keys:
        static[k].copy_(t(g[k][1:1 + B]).to(DEV))
    graph.replay()
    torch.cuda.synchronize()
    got = [p.grad.clone() for p in params]
    loss_replay, rgb_replay = loss_g.clone(), rgb_g.clone()
    assert int(m.last_raster_aux["status"].max()) == 0
    loss_e, rgb_e = step()                                    # eager on the same (new) inputs
    torch.cuda.synchronize()
    assert torch.equal(rgb_replay, rgb_e)
    np.testing.assert_allclose(float(loss_replay), float(loss_e), rtol=1e-6)
    for a, p in zip(got, params):
        ref = p.grad
        assert float((a - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-12      # atomics reorder the sums


@pytest.mark.parametrize("group,prepend", [(23, True), (1, False), (4, False)])
def test_rodrigues_kernel_matches_torch_float64(group, prepend):
    """csrc/rodrigues.cu against the torch formulation of reference utils/network_util.py:66-92 in float64, forward and backward,
    including rotation vectors at and near zero (theta -> sqrt(1e-5))."""
    from gomavatar_b200.model import rodrigues_grouped
    torch.manual_seed(group)
    n = group * 6
    r = torch.randn(n, 3) * 0.3
    r[0] = 0.0
    r[1] = 1e-6
    r[2] = torch.tensor([2.5, -1.0, 0.7])
    ro = r.double().requires_grad_(True)
    Ro = rodrigues_grouped(ro, group, prepend)                      # CPU tensors take the torch path
    g = torch.randn(Ro.shape)
    (Ro * g.double()).sum().backward()
    rk = r.to(DEV).requires_grad_(True)
    Rk = rodrigues_grouped(rk, group, prepend)
    (Rk * g.to(DEV)).sum().backward()
    assert Rk.shape == Ro.shape
    assert float((Rk.cpu().double() - Ro).abs().max()) < 2e-6
    assert float((rk.grad.cpu().double() - ro.grad).abs().max() / ro.grad.abs().max()) < 2e-5
