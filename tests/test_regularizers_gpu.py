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
    assert int(topo["row_ptr"][-1]) == topo["col"].numel() and topo["pair_vid"].shape == (conn.shape[0], 4)

    x = vb.permute(0, 2, 1).contiguous().to(DEV).requires_grad_(True)                                 # [B,3,V] like the model
    c = col.to(DEV).requires_grad_(True)
    lap, nc, cc = RG.fused_mesh_regularizers(x, c, topo)
    (10.0 * lap + 0.1 * nc + 0.05 * cc).backward()

    # torch definitions in float64
    xr = vb.double().to(DEV).requires_grad_(True)
    cr = col.double().to(DEV).requires_grad_(True)
    r_lap = RG.laplacian_smoothing(xr, faces)
    r_nc = RG.normal_consistency(xr, faces, conn)
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
