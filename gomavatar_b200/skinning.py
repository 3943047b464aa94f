"""Skeleton chain, linear-blend skinning and the Gaussians-on-mesh transform on libgom_b200.so.

Drop-in replacements (same names, argument meaning and shapes) for the reference's
``utils/body_util.py::get_global_RTs`` (:612-638) and ``apply_lbs`` (:641-644), plus ``face_gaussians`` which
replaces ``models/model.py:225-234`` (centroid, so3 exp, Steiner frame, world covariance, upper-triangle packing of
``models/modules/renderer/gaussian.py:71-75``).  All three are differentiable (hand-written backward kernels).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (GomFaceBwdArgs, GomFaceFwdArgs, GomJointBwdArgs, GomJointFwdArgs, GomLbsBwdArgs, GomLbsFwdArgs, call,
                   ptr)

# reference utils/body_util.py:36-39 (SMPL) as a parent table, parents[0] = -1
SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)
# reference utils/body_util.py SMPLX_PARENT (55 joints) — same convention
SMPLX_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15, 20, 25, 26, 20,
                 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53)

_parent_cache = {}


def _parents_tensor(parents, device):
    key = (tuple(parents), str(device))
    if key not in _parent_cache:
        _parent_cache[key] = torch.tensor(list(parents), dtype=torch.int32, device=device)
    return _parent_cache[key]


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise _lib.GomError(f"{what}: inputs must live on a CUDA device (no CPU path exists)")


def _c(t):
    return t.contiguous().float()


class _JointTransforms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cnl_gtfms, dst_Rs, dst_Ts, parents):
        _need_cuda(dst_Rs, "get_global_RTs")
        B, J = dst_Rs.shape[:2]
        cnl, Rs, Ts = _c(cnl_gtfms.detach()), _c(dst_Rs.detach()), _c(dst_Ts.detach())
        dev = Rs.device
        par = _parents_tensor(parents, dev)
        assert par.numel() == J, f"parent table has {par.numel()} joints, inputs have {J}"
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        out_R, out_T, chain, cinv = e(B, J, 3, 3), e(B, J, 3), e(B, J, 12), e(B, J, 16)
        call("gom_joint_transforms_forward", GomJointFwdArgs(
            n_frames=B, n_joints=J, parents=ptr(par), cnl_gtfms=ptr(cnl), dst_Rs=ptr(Rs), dst_Ts=ptr(Ts),
            global_Rs=ptr(out_R), global_Ts=ptr(out_T), chain_G=ptr(chain), cnl_inv=ptr(cinv)))
        ctx.save_for_backward(par, Rs, Ts, chain, cinv)
        return out_R, out_T

    @staticmethod
    def backward(ctx, gR, gT):
        par, Rs, Ts, chain, cinv = ctx.saved_tensors
        B, J = Rs.shape[:2]
        gR = torch.zeros_like(Rs) if gR is None else _c(gR)
        gT = torch.zeros_like(Ts) if gT is None else _c(gT)
        dR, dT = torch.empty_like(Rs), torch.empty_like(Ts)
        call("gom_joint_transforms_backward", GomJointBwdArgs(
            n_frames=B, n_joints=J, parents=ptr(par), dst_Rs=ptr(Rs), dst_Ts=ptr(Ts), chain_G=ptr(chain),
            cnl_inv=ptr(cinv), dL_dglobal_Rs=ptr(gR), dL_dglobal_Ts=ptr(gT), dL_ddst_Rs=ptr(dR), dL_ddst_Ts=ptr(dT)))
        return None, dR, dT, None


def get_global_RTs(cnl_gtfms, dst_Rs, dst_Ts, use_smplx=False):
    """reference utils/body_util.py:612-638.  cnl_gtfms [B,J,4,4], dst_Rs [B,J,3,3], dst_Ts [B,J,3] ->
    (scale_Rs [B,J,3,3], Ts [B,J,3]); differentiable w.r.t. dst_Rs / dst_Ts (pose refinement, train_pose.py)."""
    return _JointTransforms.apply(cnl_gtfms, dst_Rs, dst_Ts, SMPLX_PARENTS if use_smplx else SMPL_PARENTS)


class _ApplyLbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, global_Rs, global_Ts, lbs_weights):
        _need_cuda(xyz, "apply_lbs")
        B, J = global_Rs.shape[:2]
        shared = xyz.dim() == 2 or xyz.shape[0] == 1
        x = _c(xyz.detach())
        V = x.shape[-1]
        Rs, Ts, w = _c(global_Rs.detach()), _c(global_Ts.detach()), _c(lbs_weights.detach())
        assert w.shape == (J + 1, V), f"lbs_weights must be [J+1,V] = {(J + 1, V)}, got {tuple(w.shape)}"
        assert shared or x.shape[0] == B
        out = torch.empty(B, 3, V, dtype=torch.float32, device=x.device)
        call("gom_lbs_forward", GomLbsFwdArgs(
            n_frames=B, n_joints=J, n_verts=V, xyz=ptr(x), xyz_stride=0 if shared else 3 * V, lbs_weights=ptr(w),
            global_Rs=ptr(Rs), global_Ts=ptr(Ts), out=ptr(out)))
        ctx.save_for_backward(x, Rs, Ts, w)
        ctx.shared = shared
        ctx.xyz_shape = tuple(xyz.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        x, Rs, Ts, w = ctx.saved_tensors
        B, J = Rs.shape[:2]
        V = x.shape[-1]
        g = _c(g)
        pose = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = torch.empty(3, V, dtype=torch.float32, device=x.device) if ctx.shared else torch.empty_like(x)
        dR = torch.empty_like(Rs) if pose else None
        dT = torch.empty_like(Ts) if pose else None
        call("gom_lbs_backward", GomLbsBwdArgs(
            n_frames=B, n_joints=J, n_verts=V, xyz=ptr(x), xyz_stride=0 if ctx.shared else 3 * V, lbs_weights=ptr(w),
            global_Rs=ptr(Rs), global_Ts=ptr(Ts), dL_dout=ptr(g), dL_dxyz=ptr(dx),
            dL_dxyz_stride=0 if ctx.shared else 3 * V, dL_dglobal_Rs=ptr(dR), dL_dglobal_Ts=ptr(dT)))
        return dx.reshape(ctx.xyz_shape), dR, dT, None


def apply_lbs(xyzs_canonical, global_Rs, global_Ts, lbs_weights):
    """reference utils/body_util.py:641-644.  xyzs_canonical [1|B,3,V] (SoA), global_Rs [B,J,3,3], global_Ts [B,J,3],
    lbs_weights [J+1,V] (last row = background, ignored; no renormalisation)  ->  [B,3,V].
    One [1,3,V] vertex set is shared by all B frames.  Weights are a frozen buffer in the reference
    (configs/default.yaml:72-73, lr 0) and receive no gradient."""
    return _ApplyLbs.apply(xyzs_canonical, global_Rs, global_Ts, lbs_weights)


class _FaceGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces, so3, scale, sigma):
        _need_cuda(verts, "face_gaussians")
        v = _c(verts.detach())
        B, _, V = v.shape
        F = faces.shape[0]
        if faces.dtype not in (torch.int64, torch.int32):
            raise TypeError("faces must be int64 (the model's buffer) or int32")
        fc = faces.contiguous()
        w, s = _c(so3.detach()), _c(scale.detach())
        assert w.shape == (3, F) and s.shape == (3, F), "so3 / scale must be [3,F] (reference SoA layout)"
        means = torch.empty(B, F, 3, dtype=torch.float32, device=v.device)
        cov = torch.empty(B, F, 6, dtype=torch.float32, device=v.device)
        call("gom_face_gaussians_forward", GomFaceFwdArgs(
            n_frames=B, n_faces=F, n_verts=V, faces_int64=int(fc.dtype == torch.int64), sigma=float(sigma),
            verts=ptr(v), faces=ptr(fc), so3=ptr(w), scale=ptr(s), means3D=ptr(means), cov3D=ptr(cov)))
        ctx.save_for_backward(v, fc, w, s)
        ctx.sigma = float(sigma)
        return means, cov

    @staticmethod
    def backward(ctx, g_means, g_cov):
        v, fc, w, s = ctx.saved_tensors
        B, _, V = v.shape
        F = fc.shape[0]
        dev = v.device
        g_means = torch.zeros(B, F, 3, device=dev) if g_means is None else _c(g_means)
        g_cov = torch.zeros(B, F, 6, device=dev) if g_cov is None else _c(g_cov)
        dv, dw, ds = torch.empty_like(v), torch.empty_like(w), torch.empty_like(s)
        call("gom_face_gaussians_backward", GomFaceBwdArgs(
            n_frames=B, n_faces=F, n_verts=V, faces_int64=int(fc.dtype == torch.int64), sigma=ctx.sigma,
            verts=ptr(v), faces=ptr(fc), so3=ptr(w), scale=ptr(s), dL_dmeans3D=ptr(g_means), dL_dcov3D=ptr(g_cov),
            dL_dverts=ptr(dv), dL_dso3=ptr(dw), dL_dscale=ptr(ds)))
        return dv, None, dw, ds, None


def face_gaussians(vertices_observation, faces, so3, scale, sigma=1e-3):
    """reference models/model.py:225-234 (+ :27-41).  vertices_observation [B,3,V] posed vertices (SoA), faces [F,3],
    so3 / scale [3,F]  ->  (means3D [B,F,3] = face centroids, cov3D [B,F,6] = upper triangle of
    A R S S^T R^T A^T)."""
    return _FaceGaussians.apply(vertices_observation, faces, so3, scale, sigma)
