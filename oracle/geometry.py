"""Oracle (test infrastructure): torch-CPU restatement of the geometry half of the hot path.

Each function cites the reference lines it follows. dtype-generic: float32 for parity, float64 for
gradient checks.  Gradients come from torch autograd on this restatement.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def get_global_RTs(cnl_gtfms, dst_Rs, dst_Ts, parents=SMPL_PARENTS):
    """reference utils/body_util.py:612-638 (+ _construct_G_tensor :591-609).

    cnl_gtfms [B,J,4,4], dst_Rs [B,J,3,3], dst_Ts [B,J,3]  ->  Rs [B,J,3,3], Ts [B,J,3].
    Local G_i = [R_i|T_i]; chained root-to-leaf G_i = G_parent(i) G_i; F_i = G_i inv(cnl_i).
    """
    B, J = dst_Rs.shape[:2]
    local = torch.zeros(B, J, 4, 4, dtype=dst_Rs.dtype)
    local[:, :, :3, :3] = dst_Rs
    local[:, :, :3, 3] = dst_Ts
    local[:, :, 3, 3] = 1.0
    chained = [local[:, 0]]
    for i in range(1, J):
        chained.append(chained[parents[i]] @ local[:, i])
    G = torch.stack(chained, dim=1)
    Fm = G @ torch.inverse(cnl_gtfms)
    return Fm[:, :, :3, :3], Fm[:, :, :3, 3]


def apply_lbs(xyz, Rs, Ts, lbs_weights):
    """reference utils/body_util.py:641-644.  xyz [B,3,V], lbs_weights [J+1,V] (last row ignored, no
    renormalisation)  ->  [B,3,V]."""
    moved = torch.einsum("bjik,bkv->bjiv", Rs, xyz) + Ts[:, :, :, None]
    return (moved * lbs_weights[:-1][None, :, None, :]).sum(dim=1)


def rodrigues(rvec):
    """reference utils/network_util.py:66-92 (theta = sqrt(1e-5 + |r|^2)).  rvec [B,3] -> [B,3,3]."""
    theta = torch.sqrt(1e-5 + (rvec ** 2).sum(dim=1))
    r = rvec / theta[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = r[:, 0], r[:, 1], r[:, 2]
    rows = [
        x * x + (1 - x * x) * c, x * y * (1 - c) - z * s, x * z * (1 - c) + y * s,
        x * y * (1 - c) + z * s, y * y + (1 - y * y) * c, y * z * (1 - c) - x * s,
        x * z * (1 - c) - y * s, y * z * (1 - c) + x * s, z * z + (1 - z * z) * c,
    ]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def so3_exp_map(w, eps=1e-4):
    """PyTorch3D 0.7.0 ``so3_exp_map`` semantics (call site reference models/model.py:229).

    theta = sqrt(clamp(|w|^2, min=eps));  R = I + (sin th/th) K + ((1-cos th)/th^2) K^2,  K = hat(w).
    """
    nrm2 = (w * w).sum(dim=1)
    theta = torch.clamp(nrm2, min=eps).sqrt()
    inv = 1.0 / theta
    fac1 = inv * theta.sin()
    fac2 = inv * inv * (1.0 - theta.cos())
    x, y, z = w[:, 0], w[:, 1], w[:, 2]
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=1).view(-1, 3, 3)
    K2 = K @ K
    eye = torch.eye(3, dtype=w.dtype)[None]
    return fac1[:, None, None] * K + fac2[:, None, None] * K2 + eye


def steiner_frame(triangles, sigma=1e-3):
    """reference models/model.py:27-41.  triangles [F,3(vertex),3(xyz)] -> A [F,3,3] (columns 2a0|2a1|n)."""
    c = triangles.mean(dim=-2)
    f1 = 0.5 * (triangles[..., 2, :] - c)
    f2 = (1.0 / (2.0 * np.sqrt(3))) * (triangles[..., 1, :] - triangles[..., 0, :])
    t0 = torch.atan2((2 * f1 * f2).sum(-1), (f1 * f1).sum(-1) - (f2 * f2).sum(-1)) / 2
    t0 = t0[..., None]
    a0 = f1 * torch.cos(t0) + f2 * torch.sin(t0)
    a1 = f1 * torch.cos(t0 + np.pi / 2) + f2 * torch.sin(t0 + np.pi / 2)
    n = torch.cross(a0, a1, dim=-1)
    n = torch.nn.functional.normalize(n, dim=-1) * sigma
    return torch.stack([a0 * 2, a1 * 2, n], dim=-1)


def gather_triangles(verts_3V, faces):
    """reference models/model.py:225,232: verts [3,V] -> triangles [F,3,3]."""
    Fn = faces.shape[0]
    return verts_3V.permute(1, 0)[faces.reshape(-1)].reshape(Fn, 3, -1)


def face_gaussians(verts_obs_3V, faces, so3_3F, scale_3F, sigma=1e-3):
    """reference models/model.py:225-234: means [F,3] (centroids) and world covariance [F,3,3]."""
    tri = gather_triangles(verts_obs_3V, faces)
    xyz = tri.mean(dim=1)
    S = torch.diag_embed(scale_3F.permute(1, 0))
    R = so3_exp_map(so3_3F.permute(1, 0))
    cov_local = R @ S @ S.permute(0, 2, 1) @ R.permute(0, 2, 1)
    A = steiner_frame(tri, sigma)
    cov = A @ cov_local @ A.permute(0, 2, 1)
    return xyz, cov


def pack_cov6(cov):
    """reference models/modules/renderer/gaussian.py:71-75: upper triangle (xx,xy,xz,yy,yz,zz)."""
    return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], dim=-1)


def pose_geometry(vertices_3V, faces, lbs_weights, so3_3F, scale_3F, cnl_gtfms, dst_Rs, dst_Ts,
                  sigma=1e-3, global_R=None, global_T=None):
    """reference models/model.py:212-234 for ONE frame (batch index 0): returns
    (vertices_observation [3,V], means [F,3], cov [F,3,3])."""
    Rs, Ts = get_global_RTs(cnl_gtfms[None], dst_Rs[None], dst_Ts[None])
    v_obs = apply_lbs(vertices_3V[None], Rs, Ts, lbs_weights)[0]
    if global_R is not None:
        Rg = rodrigues(global_R[None])[0]
        v_obs = Rg @ v_obs + global_T[:, None]
    xyz, cov = face_gaussians(v_obs, faces, so3_3F, scale_3F, sigma)
    return v_obs, xyz, cov
