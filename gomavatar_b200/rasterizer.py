"""Differentiable tile-based 3D-Gaussian splat rasterizer on libgom_b200.so, behind the API surface of the
third-party ``diff_gaussian_rasterization`` package that the reference imports at
``models/modules/renderer/gaussian.py:9`` (``GaussianRasterizationSettings`` / ``GaussianRasterizer``;
SURVEY.md §8b-2, App. A.1), plus the batched, fused entry point the B200 hot path uses (``rasterize_gaussians``).

Differences from upstream that are visible to a caller: none in results (see tests); internally there is no
``num_rendered`` read-back — the kernels use fixed-capacity instance buffers and flag overflow on the device.
``strict=True`` (the drop-in default) reads that flag right after the launch (one sync, like upstream) and re-runs
with a larger buffer; ``strict=False`` (training loop) leaves the check to the caller (``aux['status']``).
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import GomRasterBwdArgs, GomRasterFwdArgs, check, ptr

TILE = 16


def default_capacity(n_gauss: int) -> int:
    """Per-frame (Gaussian, tile) instance capacity: generous for mesh-attached Gaussians (N_dup ~ 2-3 P)."""
    return max(16 * int(n_gauss), 1 << 16)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    return t.detach().contiguous().float()


class _Rasterize(torch.autograd.Function):
    """inputs batched: means3D [B,P,3], cov3D [B,P,6], colors [P,C] | [B,P,C], opacities [B,P],
    means2D None | [B,P,3] (gradient sink only, as upstream's screenspace_points)."""

    @staticmethod
    def forward(ctx, means3D, cov3D, colors, opacities, means2D, view, proj, tanfov, bg, H, W, interleaved, strict,
                capacity, aux, color_grad_channels=0):
        L = _lib.lib()
        dev = means3D.device
        if dev.type != "cuda":
            raise _lib.GomError("rasterize_gaussians: inputs must live on a CUDA device (no CPU path exists)")
        B, P, _ = means3D.shape
        shared_colors = colors.dim() == 2
        C = colors.shape[-1]
        m3, c3, col, op = _f32c(means3D), _f32c(cov3D), _f32c(colors), _f32c(opacities)
        view, proj, tanfov, bg = _f32c(view).reshape(B, 16), _f32c(proj).reshape(B, 16), _f32c(tanfov).reshape(B, 2), _f32c(bg)
        bg = bg.reshape(B, -1)[:, :C].contiguous()
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        T = gx * gy
        cap = int(capacity) if capacity else default_capacity(P)
        e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        out_shape = (B, H, W, C) if interleaved else (B, C, H, W)
        while True:
            st = dict(
                out_color=e(*out_shape), final_T=e(B, H, W), n_contrib=e(B, H, W, dtype=torch.int32),
                radii=e(B, P, dtype=torch.int32), depth=e(B, P), xy=e(B, P, 2), conic_opacity=e(B, P, 4),
                rect=e(B, P, 4, dtype=torch.int32), tile_count=e(B, T, dtype=torch.int32),
                tile_offset=e(B, T + 1, dtype=torch.int32), tile_cursor=e(B, T, dtype=torch.int32),
                inst_keys=e(B, cap, dtype=torch.int64), point_list=e(B, cap, dtype=torch.int32),
                status=e(B, dtype=torch.int32), worklist=e(B * T, dtype=torch.int32),
                point_mask=e(B, cap, dtype=torch.uint8))
            a = GomRasterFwdArgs(
                n_frames=B, n_gauss=P, height=H, width=W, n_channels=C, interleaved=int(bool(interleaved)),
                inst_capacity=cap,
                means3D=ptr(m3), means3D_stride=P * 3, cov3D=ptr(c3), cov3D_stride=P * 6,
                colors=ptr(col), colors_stride=0 if shared_colors else P * C,
                opacities=ptr(op), opacities_stride=P,
                viewmatrix=ptr(view), projmatrix=ptr(proj), tanfov=ptr(tanfov), bg=ptr(bg),
                **{k: ptr(v) for k, v in st.items()})
            check(L.gom_raster_forward(ctypes.byref(a), _stream()), "gom_raster_forward")
            if strict and int(st["status"].max().item()) & _lib.STATUS_OVERFLOW:     # one sync, like upstream
                need = int(st["tile_offset"][:, T].to(torch.int64).bitwise_and(0xFFFFFFFF).max().item())
                cap = int(need * 1.25) + 1024
                continue
            break
        ctx.dims = (B, P, H, W, C, bool(interleaved), cap, shared_colors)
        ctx.color_grad_channels = int(color_grad_channels or 0)
        ctx.has_means2D = means2D is not None
        ctx.save_for_backward(m3, c3, col, view, proj, tanfov, bg, st["final_T"], st["n_contrib"], st["radii"],
                              st["xy"], st["conic_opacity"], st["tile_offset"], st["point_list"], st["worklist"], st["point_mask"])
        if aux is not None:
            aux.update(st)
            aux["inst_capacity"] = cap
        ctx.mark_non_differentiable(st["radii"], st["final_T"], st["n_contrib"])
        return st["out_color"], st["radii"], st["final_T"], st["n_contrib"]

    @staticmethod
    def backward(ctx, g_color, _g_radii, _g_T, _g_n):
        L = _lib.lib()
        B, P, H, W, C, interleaved, cap, shared_colors = ctx.dims
        (m3, c3, col, view, proj, tanfov, bg, final_T, n_contrib, radii, xy, conic_opacity, tile_offset,
         point_list, worklist, point_mask) = ctx.saved_tensors
        dev = m3.device
        g_color = g_color.contiguous().float()
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        d_means3D, d_cov3D = e(B, P, 3), e(B, P, 6)
        d_colors = e(P, C) if shared_colors else e(B, P, C)
        need_op = ctx.needs_input_grad[3]
        d_op = e(B, P) if need_op else None
        d_mean2D, d_conic = e(B, P, 2), e(B, P, 3)
        a = GomRasterBwdArgs(
            n_frames=B, n_gauss=P, height=H, width=W, n_channels=C, interleaved=int(interleaved), inst_capacity=cap,
            color_grad_channels=ctx.color_grad_channels,
            means3D=ptr(m3), means3D_stride=P * 3, cov3D=ptr(c3), cov3D_stride=P * 6,
            colors=ptr(col), colors_stride=0 if shared_colors else P * C,
            viewmatrix=ptr(view), projmatrix=ptr(proj), tanfov=ptr(tanfov), bg=ptr(bg),
            final_T=ptr(final_T), n_contrib=ptr(n_contrib), radii=ptr(radii), xy=ptr(xy),
            conic_opacity=ptr(conic_opacity), tile_offset=ptr(tile_offset), point_list=ptr(point_list),
            worklist=ptr(worklist), point_mask=ptr(point_mask), dL_dout=ptr(g_color), dL_dmeans3D=ptr(d_means3D), dL_dcov3D=ptr(d_cov3D),
            dL_dcolors=ptr(d_colors), dL_dcolors_stride=0 if shared_colors else P * C,
            dL_dopacity=ptr(d_op), dL_dmeans2D=ptr(d_mean2D), dL_dconic=ptr(d_conic))
        check(L.gom_raster_backward(ctypes.byref(a), _stream()), "gom_raster_backward")
        g_means2D = None
        if ctx.has_means2D and ctx.needs_input_grad[4]:
            g_means2D = torch.cat([d_mean2D, torch.zeros_like(d_mean2D[..., :1])], dim=-1)
        return (d_means3D, d_cov3D, d_colors, d_op, g_means2D) + (None,) * 11


def rasterize_gaussians(means3D, cov3D, colors, opacities, viewmatrix, projmatrix, tanfov, bg, image_height,
                        image_width, means2D=None, interleaved=False, strict=True, capacity=None, aux=None,
                        color_grad_channels=0):
    """Batched differentiable splatting of B frames in one launch sequence.

    means3D [B,P,3], cov3D [B,P,6] (xx,xy,xz,yy,yz,zz), colors [P,C] (shared by all frames) or [B,P,C] with C in
    {3,4}, opacities [B,P]; viewmatrix/projmatrix [B,4,4] exactly as the reference builds them (E^T, E^T K_ndc^T);
    tanfov [B,2]; bg [B,C].  Returns (color [B,C,H,W] or [B,H,W,C] if interleaved, radii [B,P] int32,
    final_T [B,H,W], n_contrib [B,H,W]).  ``aux`` (a dict) receives every intermediate buffer.
    ``color_grad_channels=3`` with 4-channel colours skips the gradient of the 4th channel (GoMAvatar renders alpha with
    a constant-1 "colour", reference gaussian.py:49), which puts the backward on its 8-component fast path.
    """
    return _Rasterize.apply(means3D, cov3D, colors, opacities, means2D, viewmatrix, projmatrix, tanfov, bg,
                            int(image_height), int(image_width), bool(interleaved), bool(strict), capacity, aux,
                            int(color_grad_channels or 0))


# ----------------------------------------------------------------------------------------------------------------
# The reference-facing API (same names, argument meaning and error behaviour as diff_gaussian_rasterization)
# ----------------------------------------------------------------------------------------------------------------
class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    """Drop-in for ``diff_gaussian_rasterization.GaussianRasterizer`` as used at reference gaussian.py:20,67,83-91."""

    def __init__(self, raster_settings: Optional[GaussianRasterizationSettings]):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        s = self.raster_settings
        with torch.no_grad():
            view = s.viewmatrix.contiguous().float()
            z = positions.float() @ view[:3, 2] + view[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        s = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is not None or cov3D_precomp is None:
            raise NotImplementedError(
                "gomavatar_b200 implements the branch GoMAvatar uses: colors_precomp + cov3D_precomp "
                "(reference models/modules/renderer/gaussian.py:83-91); SH / scale-rotation inputs are out of scope")
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if s.scale_modifier != 1.0:
            raise NotImplementedError("scale_modifier != 1 is not used by the reference with cov3D_precomp")
        C = colors_precomp.shape[1]
        dev = means3D.device
        tanfov = torch.tensor([[s.tanfovx, s.tanfovy]], dtype=torch.float32, device=dev)
        color, radii, _, _ = rasterize_gaussians(
            means3D[None], cov3D_precomp[None], colors_precomp, opacities.reshape(1, -1),
            s.viewmatrix.reshape(1, 4, 4), s.projmatrix.reshape(1, 4, 4), tanfov, s.bg.reshape(1, -1)[:, :C],
            s.image_height, s.image_width, means2D=None if means2D is None else means2D[None], strict=True)
        return color[0], radii[0]
