"""GPU: the fused mesh regularisers (csrc/regularizers.cu through gom_mesh_regularizers) against the torch definitions of
gomavatar_b200/regularizers.py (which tests/test_regularizers_cpu.py checks against dense restatements) — values and the
gradients w.r.t. the vertices of every mesh of the batch and the face colours."""
import numpy as np
import pytest
import torch

from gomavatar_b200 import regularizers as RG
from gomavatar_b200 import synthetic as S
from gomavatar_b200.model import mesh_edges

pytestmark = pytest.mark.gpu
DEV = "cuda"
t = torch.from_numpy


@pytest.mark.parametrize("n_faces,B", [(2000, 3), (13776, 1), (30000, 2)])
def test_fused_regularisers_match_torch_definitions(n_faces, B):
    sc = S.make_humanoid(n_faces, seed=0)
    g = torch.Generator().manual_seed(1)
    v = t(sc.vertices).float()
    vb = torch.stack([v + 0.004 * torch.randn(v.shape, generator=g) for _ in range(B)])            # [B,V,3]
    _, conn = mesh_edges(sc.faces.astype(np.int64), sc.vertices)
    faces, conn = t(sc.faces).long().to(DEV), t(conn).to(DEV)
    col = torch.rand(sc.n_faces, 3, generator=g)
    V = v.shape[0]
    topo = RG.mesh_topology(faces, conn, V)
    assert int(topo["row_ptr"][-1]) == topo["col"].numel()
    # normal term: every edge-sharing pair (PyTorch3D); colour term: the model's connectivity (one pair short, model.py:119-123)
    assert topo["pair_vid"].shape == (conn.shape[0] + 1, 4) and topo["pair_face"].shape == (conn.shape[0], 2)

    x = vb.permute(0, 2, 1).contiguous().to(DEV).requires_grad_(True)                                 # [B,3,V] like the model
    c = col.to(DEV).requires_grad_(True)
    lap, nc, cc = RG.fused_mesh_regularizers(x, c, topo)
    (10.0 * lap + 0.1 * nc + 0.05 * cc).backward()

    # torch definitions in float64
    xr = vb.double().to(DEV).requires_grad_(True)
    cr = col.double().to(DEV).requires_grad_(True)
    r_lap = RG.laplacian_smoothing(xr, faces)
    r_nc = RG.normal_consistency(xr, faces)
    r_cc = RG.color_consistency(cr, conn)
    (10.0 * r_lap + 0.1 * r_nc + 0.05 * r_cc).backward()
    for got, ref, what in ((lap, r_lap, "laplacian"), (nc, r_nc, "normal"), (cc, r_cc, "colour")):
        assert abs(float(got) - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-9, (what, float(got), float(ref))
    gx, rx = x.grad.permute(0, 2, 1).double(), xr.grad
    assert float((gx - rx).abs().max()) <= 1e-3 * float(rx.abs().max()), float((gx - rx).abs().max() / rx.abs().max())
    assert float((c.grad.double() - cr.grad).abs().max()) <= 1e-6 * float(cr.grad.abs().max()) + 1e-12


def test_single_terms_and_compute_loss_route():
    """each term can be switched off; compute_loss takes the fused route on CUDA tensors and equals the torch route"""
    sc = S.make_humanoid(2000, seed=0)
    _, conn = mesh_edges(sc.faces.astype(np.int64), sc.vertices)
    faces, conn = t(sc.faces).long().to(DEV), t(conn).to(DEV)
    v = t(sc.vertices).float().to(DEV)
    x = (v + 0.003 * torch.randn_like(v)).t()[None].contiguous()
    col = torch.rand(sc.n_faces, 3, device=DEV)
    topo = RG.mesh_topology(faces, conn, v.shape[0])
    lap, nc, cc = RG.fused_mesh_regularizers(x, col, topo, laplacian=True, normal=False, color=False)
    assert float(nc) == 0.0 and float(cc) == 0.0
    ref = RG.laplacian_smoothing(x[0].t().double(), faces)
    assert abs(float(lap) - float(ref)) <= 2e-5 * float(ref)


def test_compute_loss_equals_the_reference_compute_loss(golden_dir):
    """regularizers.compute_loss (unpack + L1 rgb / mask kernels + LPIPS + fused regulariser kernels + dilated normal-mask L1)
    against the reference's own unpack + compute_loss (train.py:53-55, :98-163) run by oracle/make_golden.py::loss_golden
    with the coefficients of exps/zju-mocap_377.yaml: every term, its scaling and the total."""
    import os
    import types
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    g = np.load(os.path.join(golden_dir, "golden_loss.npz"))
    heads = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    lp = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)], conv_precision="fp32").to(DEV)
    d = lambda k: t(g[k]).to(DEV)
    verts = d("verts").float()
    model = types.SimpleNamespace(faces=d("faces").long(), vertices=verts.t().contiguous())
    outputs = {"vertices_observation": verts.t()[None].contiguous(), "colors": d("colors").float(),
               "face_connectivity": d("face_connectivity").long(), "normal_mask": d("normal_mask").float()}
    cfg = {"rgb": {"coeff": 1.0}, "mask": {"coeff": 5.0}, "lpips": {"coeff": 1.0},
           "laplacian": {"coeff_canonical": 0.0, "coeff_observation": 10.0},
           "normal": {"mask_dilate": True, "kernel_size": 7, "coeff_mask": 1.0, "coeff_consist": 0.10}, "color_consist": {"coeff": 0.050}}
    total, losses = RG.compute_loss(d("rgb_raw").float(), d("mask_pred").float(), d("bgcolor").float(), d("rgb_gt").float(),
                                    d("mask_gt").float(), outputs, model, cfg, lpips_func=lp)
    names = {k[len("unscaled."):] for k in g.files if k.startswith("unscaled.")}
    assert set(losses) == names, set(losses) ^ names
    for k in names:
        for kind in ("unscaled", "scaled"):
            ref, got = float(g[f"{kind}.{k}"]), float(losses[k][kind])
            tol = 2e-4 if k == "lpips" else 2e-5                      # LPIPS: different convolution implementations
            assert abs(got - ref) <= tol * abs(ref) + 1e-8, (k, kind, got, ref)
    assert abs(float(total) - float(g["total"])) <= 5e-5 * float(g["total"])
