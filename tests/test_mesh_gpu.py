"""GPU parity: csrc/mesh_raster.cu (normal map + soft silhouette, forward / backward) against the torch restatement of
PyTorch3D's rasterizer semantics in oracle/mesh_raster.py (parity unpinned: pytorch3d is absent), and the
reference-shaped ``mesh_renderer.Renderer`` end to end (reference models/modules/renderer/mesh.py:64-128)."""
import math

import numpy as np
import pytest
import torch

from gomavatar_b200 import synthetic as S
from oracle import geometry as G
from oracle import mesh_raster as MR

pytestmark = pytest.mark.gpu
t = torch.from_numpy
DEV = "cuda:0"


def _posed_scene(n_faces, size, seed=3, focal=537.0, distance=3.5):
    W, H = size
    sc = S.make_humanoid(n_faces, seed=0)
    fr = S.make_frames(sc, 1, img_size=(W, H), seed=seed, focal=focal * W / 512.0, distance=distance, base_size=W)
    pr = S.make_params(sc, seed=1)
    v_obs, _, _ = G.pose_geometry(t(pr["vertices"]), t(sc.faces), t(sc.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                  t(fr["cnl_gtfms"][0]), t(fr["dst_Rs"][0]), t(fr["dst_Ts"][0]))
    return sc, fr, v_obs.detach().T.contiguous()           # posed vertices [V,3]


@pytest.mark.parametrize("n_faces,size,sigma,K", [(2000, (64, 64), 1e-5, 50), (2000, (80, 56), 1e-4, 50), (4000, (100, 120), 1e-4, 50),
                                                  (2000, (64, 64), 1e-3, 4)])
def test_rasterize_mesh_forward_backward_matches_oracle(n_faces, size, sigma, K):
    from gomavatar_b200.mesh_renderer import rasterize_mesh
    W, H = size
    sc, fr, verts = _posed_scene(n_faces, size)
    faces = t(sc.faces).long()
    Kc, E = t(fr["K"][0]), t(fr["E"][0])
    blur = math.log(1. / 1e-4 - 1.) * sigma
    # oracle (float64 for a clean reference of the gradients)
    ndc_o = MR.ndc_T_world(verts.double(), Kc.double(), E.double(), H, W).detach().requires_grad_(True)
    vn_o = (MR.vertex_normals(verts.double(), faces) @ E[:3, :3].double().T).detach().requires_grad_(True)
    p2f_h, _, _ = MR.rasterize(ndc_o, faces, H, W, 0.0, 1)
    nm_o = MR.normal_map(p2f_h, faces, vn_o)
    p2f_s, _, d_s = MR.rasterize(ndc_o, faces, H, W, blur, K)
    al_o = MR.soft_silhouette(p2f_s, d_s)
    rng = np.random.default_rng(7)
    g_n, g_a = rng.normal(size=(H, W, 3)).astype(np.float32), rng.normal(size=(H, W)).astype(np.float32)
    ((nm_o * t(g_n).double()).sum() + (al_o * t(g_a).double()).sum()).backward()
    # kernels
    ndc_k = ndc_o.detach().float()[None].to(DEV).requires_grad_(True)
    vn_k = vn_o.detach().float()[None].to(DEV).requires_grad_(True)
    aux = {}
    nm_k, al_k, p2f_k = rasterize_mesh(ndc_k, vn_k, faces.to(DEV), H, W, soft=True, blur_radius=blur, faces_per_pixel=K, aux=aux)
    assert int(aux["status"][0]) == 0
    hit_o = p2f_h[..., 0].numpy()
    hit_k = p2f_k[0].cpu().numpy()
    assert (hit_o >= 0).mean() > 0.03, "the subject must be in view"
    same = hit_o == hit_k
    assert (~same).mean() <= 2e-3, f"pix_to_face differs on {(~same).mean():.2%} of the pixels"      # fp32 edge-function ties
    np.testing.assert_allclose(nm_k[0].detach().cpu().numpy()[same], nm_o.detach().float().numpy()[same], atol=2e-6)
    da = np.abs(al_k[0].detach().cpu().numpy() - al_o.detach().float().numpy())
    assert (da > 2e-3).mean() <= 2e-3 and np.quantile(da, 0.99) < 2e-4, (float(da.max()), float((da > 2e-3).mean()))
    if K < 50:
        assert float(torch.isfinite(aux["zcut"]).float().mean()) > 0.01, "the K-nearest slow path must be exercised"
    ((nm_k * t(g_n).to(DEV)).sum() + (al_k * t(g_a).to(DEV)).sum()).backward()
    gv_o, gv_k = ndc_o.grad.float().numpy(), ndc_k.grad[0].cpu().numpy()
    assert float(np.abs(gv_k[:, 2]).max()) == 0.0
    err = np.abs(gv_k - gv_o) / np.abs(gv_o).max()
    assert (err > 1e-3).mean() <= 5e-3 and np.sqrt((err ** 2).sum() / ((gv_o / np.abs(gv_o).max()) ** 2).sum()) < 3e-2, \
        (float(err.max()), float((err > 1e-3).mean()))
    gn_o, gn_k = vn_o.grad.float().numpy(), vn_k.grad[0].cpu().numpy()
    errn = np.abs(gn_k - gn_o) / np.abs(gn_o).max()
    assert (errn > 1e-3).mean() <= 5e-3, float((errn > 1e-3).mean())


@pytest.mark.parametrize("training", [True, False])
def test_renderer_module_matches_oracle_render(training):
    """reference call pattern (models/model.py:271-274): world-space posed vertices [B,3,V], camera-space vertex normals."""
    from gomavatar_b200.mesh_renderer import Renderer, vertex_normals
    W, H = 96, 96
    sc, fr, verts = _posed_scene(3000, (W, H))
    faces = t(sc.faces).long()
    r = Renderer({"img_size": [W, H], "sigma": 1e-5}).to(DEV)
    r.train(training)
    xyz = verts.T[None].to(DEV).requires_grad_(True)                 # [1,3,V]
    Kd, Ed = t(fr["K"][:1]).to(DEV), t(fr["E"][:1]).to(DEV)
    vn = vertex_normals(xyz.permute(0, 2, 1), faces.to(DEV))
    vn = torch.bmm(Ed[:, :3, :3], vn.permute(0, 2, 1)).permute(0, 2, 1)
    normal, mask = r(xyz, vn, Kd, Ed, faces=faces.to(DEV))
    vo = verts.double().requires_grad_(True)
    nm_o, m_o = MR.render(vo, faces, t(fr["K"][0]).double(), t(fr["E"][0]).double(), H, W, training=training, sigma_cfg=1e-5)
    d = np.abs(normal[0].detach().cpu().numpy() - nm_o.detach().float().numpy()).max(-1)
    assert (d > 1e-4).mean() <= 2e-3
    if not training:
        assert mask is None
        return
    assert mask.shape == (1, H, W, 1)
    da = np.abs(mask[0, ..., 0].detach().cpu().numpy() - m_o.detach().float().numpy())
    assert (da > 2e-3).mean() <= 2e-3
    rng = np.random.default_rng(1)
    g_a = rng.normal(size=(H, W)).astype(np.float32)
    g_n = rng.normal(size=(H, W, 3)).astype(np.float32)
    ((mask[0, ..., 0] * t(g_a).to(DEV)).sum() + (normal[0] * t(g_n).to(DEV)).sum()).backward()
    ((m_o * t(g_a).double()).sum() + (nm_o * t(g_n).double()).sum()).backward()
    go, gk = vo.grad.float().numpy(), xyz.grad[0].T.cpu().numpy()
    rel_l2 = np.sqrt(((gk - go) ** 2).sum() / (go ** 2).sum())
    assert rel_l2 < 3e-2, rel_l2


@pytest.mark.parametrize("size", [(96, 96), (80, 56), (56, 80)])
def test_ndc_and_vertex_normal_kernels_match_torch_float64(size):
    """csrc/mesh_prep.cu: ndc_T_world (reference utils/pc_util.py:30-46) and camera-space vertex normals (reference
    models/model.py:271-273 on PyTorch3D's verts_normals_padded), forward and backward, against the torch formulations of
    gomavatar_b200.mesh_renderer evaluated in float64."""
    from gomavatar_b200.mesh_renderer import _NdcTWorld, ndc_T_world, vertex_normals, vertex_normals_cam
    W, H = size
    sc = S.make_humanoid(3000, seed=0)
    fr = S.make_frames(sc, 3, img_size=(W, H), seed=5, focal=537.0 * W / 512.0, base_size=W)
    pr = S.make_params(sc, seed=1)
    faces = t(sc.faces).long()
    vs = []
    for b in range(3):
        v, _, _ = G.pose_geometry(t(pr["vertices"]), faces, t(sc.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                  t(fr["cnl_gtfms"][b]), t(fr["dst_Rs"][b]), t(fr["dst_Ts"][b]))
        vs.append(v.detach())
    xyz = torch.stack(vs)                                                   # [B,3,V]
    Kd, Ed = t(fr["K"]), t(fr["E"])
    rng = np.random.default_rng(2)
    g_ndc = t(rng.normal(size=(3, xyz.shape[2], 3)).astype(np.float32))
    g_vn = t(rng.normal(size=(3, xyz.shape[2], 3)).astype(np.float32))
    # float64 torch reference
    xo = xyz.double().requires_grad_(True)
    ndc_o = ndc_T_world(xo, Kd.double(), Ed.double(), H, W)
    n_o = vertex_normals(xo.permute(0, 2, 1), faces)
    n_o = torch.bmm(Ed[:, :3, :3].double(), n_o.permute(0, 2, 1)).permute(0, 2, 1)
    ((ndc_o * g_ndc.double()).sum() + (n_o * g_vn.double()).sum()).backward()
    # kernels, separately so that each gradient is checked on its own
    xk = xyz.to(DEV).requires_grad_(True)
    ndc_k = _NdcTWorld.apply(xk, Kd.to(DEV), Ed.to(DEV), H, W)
    (ndc_k * g_ndc.to(DEV)).sum().backward()
    g1 = xk.grad.clone(); xk.grad = None
    n_k = vertex_normals_cam(xk, faces.to(DEV), Ed.to(DEV))
    (n_k * g_vn.to(DEV)).sum().backward()
    g2 = xk.grad.clone()
    assert float((ndc_k.cpu().double() - ndc_o).abs().max()) < 2e-5
    assert float((n_k.cpu().double() - n_o).abs().max()) < 2e-5
    xo2 = xyz.double().requires_grad_(True)
    (ndc_T_world(xo2, Kd.double(), Ed.double(), H, W) * g_ndc.double()).sum().backward()
    g1_o = xo2.grad
    g2_o = xo.grad - g1_o
    assert float((g1.cpu().double() - g1_o).abs().max() / g1_o.abs().max()) < 2e-5
    assert float((g2.cpu().double() - g2_o).abs().max() / g2_o.abs().max()) < 1e-4


@pytest.mark.parametrize("B,H,W,k,dilate", [(2, 64, 64, 7, True), (1, 50, 70, 7, True), (3, 33, 40, 3, True), (2, 64, 48, 7, False),
                                            (1, 40, 40, 15, True)])
def test_dilated_mask_l1_kernel_matches_torch(B, H, W, k, dilate):
    """csrc/mesh_prep.cu gom_dilated_mask_l1 against reference train.py:137-146 written with F.max_pool2d."""
    import torch.nn.functional as F
    from gomavatar_b200 import regularizers as RG
    g = torch.Generator().manual_seed(B * H + k)
    gt = (torch.rand(B, H, W, generator=g) > 0.7).float()
    nm = torch.rand(B, H, W, generator=g)
    nm[:, ::5, ::3] = 1.0                                                     # exact ties: |0| has gradient 0 in torch
    ref_in = nm.clone().double().requires_grad_(True)
    d = F.max_pool2d(gt.double().unsqueeze(1), kernel_size=k, stride=1, padding=k // 2).squeeze(1) if dilate else gt.double()
    ref = (ref_in - d).abs().mean()
    ref.backward()
    x = nm.to(DEV).requires_grad_(True)
    out = RG.normal_mask_loss(x, gt.to(DEV), k, dilate)
    (out * 3.0).backward()
    assert abs(float(out) - float(ref)) < 1e-6
    assert float((x.grad.cpu().double() - 3.0 * ref_in.grad).abs().max()) < 1e-9 + 1e-6 * float(ref_in.grad.abs().max())


@pytest.mark.parametrize("n_faces,size,K", [(30000, (256, 256), 50), (12000, (160, 160), 50), (12000, (160, 160), 8), (120000, (512, 512), 50)])
def test_tile_kernels_equal_one_block_per_tile_kernel_on_dense_meshes(n_faces, size, K, monkeypatch):
    """The worklist / 8-slice tile kernels (k_mesh_tiles_fwd/bwd) against round 1's one-block-per-tile forward kernel
    (GOM_MESH_LEGACY=1, itself checked against the oracle above) where the K-nearest selection matters: faces much smaller than
    a pixel, 50 - 400 soft candidates per pixel, most body pixels beyond K.  pix_to_face identical, alpha to rounding, the same
    pixels carry a cut, with bit-identical cut depths / ids (the kernels share one definition of the per-(pixel, face) arithmetic, written with
    explicit rounding intrinsics so that nvcc cannot contract it differently in different kernels: at 120 000 faces plain
    expressions made 1 cut pixel in 16 000 differ;
    the backward replays the cut by comparing recomputed depths with the stored one)."""
    import os
    from gomavatar_b200.mesh_renderer import _NdcTWorld, rasterize_mesh, vertex_normals_cam
    W, H = size
    sc, fr, verts = _posed_scene(n_faces, size)
    faces = t(sc.faces).long().to(DEV)
    xyz = verts.T[None].contiguous().to(DEV)
    Kd, Ed = t(fr["K"][:1]).to(DEV), t(fr["E"][:1]).to(DEV)
    ndc = _NdcTWorld.apply(xyz, Kd, Ed, H, W)
    vn = vertex_normals_cam(xyz, faces, Ed)
    blur = math.log(1. / 1e-4 - 1.) * 1e-5
    rng = np.random.default_rng(0)
    g_n, g_a = t(rng.normal(size=(1, H, W, 3)).astype(np.float32)).to(DEV), t(rng.normal(size=(1, H, W)).astype(np.float32)).to(DEV)
    res = {}
    for legacy in ("1", "0"):
        monkeypatch.setenv("GOM_MESH_LEGACY", legacy)
        a, b = ndc.detach().clone().requires_grad_(True), vn.detach().clone().requires_grad_(True)
        aux = {}
        nm, al, p2f = rasterize_mesh(a, b, faces, H, W, soft=True, blur_radius=blur, faces_per_pixel=K, aux=aux, capacity=32 * n_faces)
        assert int(aux["status"][0]) == 0
        torch.autograd.backward([nm, al], [g_n, g_a])
        res[legacy] = (nm.detach(), al.detach(), p2f, aux["zcut"].clone(), aux["idcut"].clone(), a.grad.clone(), b.grad.clone())
    monkeypatch.delenv("GOM_MESH_LEGACY")
    (nm0, al0, p0, z0, i0, gv0, gn0), (nm1, al1, p1, z1, i1, gv1, gn1) = res["1"], res["0"]
    assert float((p0 >= 0).float().mean()) > 0.02
    assert bool((p0 == p1).all())
    assert float((nm0 - nm1).abs().max()) == 0.0
    assert float((al0 - al1).abs().max()) < 2e-6
    cut0, cut1 = torch.isfinite(z0), torch.isfinite(z1)
    assert float(cut0.float().mean()) > 0.005, "the K-nearest selection must be exercised"
    assert bool((cut0 == cut1).all())
    # every kernel of csrc/mesh_raster.cu shares ONE definition of the per-(pixel, face) arithmetic: the cuts are bit-identical
    assert bool((z0[cut0] == z1[cut0]).all())
    assert bool((i0[cut0] == i1[cut0]).all())
    worst = float((gv0 - gv1).abs().max() / gv0.abs().max())
    rel_l2 = float((gv0 - gv1).norm() / gv0.norm())
    assert worst < 1e-4 and rel_l2 < 1e-5, (worst, rel_l2)          # same cut, same candidates: only the order of the float atomics differs
    assert float((gn0 - gn1).abs().max() / gn0.abs().max()) < 1e-5
