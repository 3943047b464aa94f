"""Photometric losses of the training step on libgom_b200.so.

``unpack``          drop-in for reference train.py:53-55 (background compositing with a per-frame colour)
``photometric_l1``  unpack + L1(rgb) + L1(mask) of train.py:101-111 in one forward and one backward kernel; returns the
                    composited image too, because LPIPS (train.py:113-121) consumes it
``compute_loss``    the three photometric terms of reference ``train.py::compute_loss`` with its coefficients
                    (configs/default.yaml:101-106: rgb 1.0, mask 5.0, lpips 1.0); the mesh regularisers
                    (train.py:123-160) are out of the hot path and stay with the caller.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomPhotoArgs, GomShadeArgs, call, ptr


def _pixel_view(t, channels):
    """(tensor keeping the storage alive, pixel stride) for a [B,H,W,(C)] tensor that is either contiguous or a channel
    slice of a contiguous interleaved [B,H,W,K] tensor (e.g. rgba[..., :3], rgba[..., 3])."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    B, H, W = t.shape[:3]
    st = t.stride()
    ps = st[2]
    ok = (channels == 1 or st[3] == 1) and st[1] == W * ps and st[0] == H * W * ps and ps >= channels
    if not ok:
        t = t.contiguous()
        ps = channels
    return t, ps


class _PhotometricL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgbs, masks, bgcolors, rgb_gt, mask_gt):
        if rgbs.device.type != "cuda":
            raise _lib.GomError("photometric_l1: inputs must live on a CUDA device (no CPU path exists)")
        B, H, W, _ = rgbs.shape
        rgb, rgb_ps = _pixel_view(rgbs, 3)
        mask, mask_ps = _pixel_view(masks, 1)
        bg = None if bgcolors is None else bgcolors.detach().contiguous().float()
        gt_rgb = None if rgb_gt is None else rgb_gt.detach().contiguous().float()
        gt_mask = None if mask_gt is None else mask_gt.detach().contiguous().float()
        dev = rgbs.device
        unpacked = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev)
        sums = torch.empty(2, dtype=torch.float32, device=dev)
        call("gom_photometric_forward", GomPhotoArgs(
            n_frames=B, height=H, width=W, rgb=ptr(rgb), rgb_pixel_stride=rgb_ps, mask=ptr(mask), mask_pixel_stride=mask_ps,
            bgcolor=ptr(bg), gt_rgb=ptr(gt_rgb), gt_mask=ptr(gt_mask), unpacked=ptr(unpacked), loss_sums=ptr(sums)))
        ctx.save_for_backward(rgb, mask, bg, gt_rgb, gt_mask)
        ctx.meta = (B, H, W, rgb_ps, mask_ps)
        n = float(B * H * W)
        return unpacked, sums[0] / (3.0 * n), sums[1] / n

    @staticmethod
    def backward(ctx, g_unpacked, g_lrgb, g_lmask):
        rgb, mask, bg, gt_rgb, gt_mask = ctx.saved_tensors
        B, H, W, rgb_ps, mask_ps = ctx.meta
        dev = rgb.device
        z = torch.zeros((), device=dev)
        g_loss = torch.stack([z if g_lrgb is None else g_lrgb.float(), z if g_lmask is None else g_lmask.float()])
        g_u = None if g_unpacked is None else g_unpacked.contiguous().float()
        d_rgb = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev)
        d_mask = torch.empty(B, H, W, dtype=torch.float32, device=dev)
        call("gom_photometric_backward", GomPhotoArgs(
            n_frames=B, height=H, width=W, rgb=ptr(rgb), rgb_pixel_stride=rgb_ps, mask=ptr(mask), mask_pixel_stride=mask_ps,
            bgcolor=ptr(bg), gt_rgb=ptr(gt_rgb), gt_mask=ptr(gt_mask), dL_dunpacked=ptr(g_u), dL_dlosses=ptr(g_loss),
            dL_drgb=ptr(d_rgb), dL_drgb_pixel_stride=3, dL_dmask=ptr(d_mask), dL_dmask_pixel_stride=1))
        return d_rgb, d_mask, None, None, None


class _PhotometricL1RGBA(torch.autograd.Function):
    """The same two kernels on the rasterizer's interleaved [B,H,W,4] output itself: the gradient is written as ONE [B,H,W,4]
    tensor (pixel stride 4 for both parts).  Autograd would otherwise rebuild it from the gradients of the two channel slices
    with two zero fills, two strided copies and an add over the full image (~50 us per 8 frames of 512x512)."""

    @staticmethod
    def forward(ctx, rgba, bgcolors, rgb_gt, mask_gt):
        B, H, W, _ = rgba.shape
        x = rgba.detach()
        bg = None if bgcolors is None else bgcolors.detach().contiguous().float()
        gt_rgb = None if rgb_gt is None else rgb_gt.detach().contiguous().float()
        gt_mask = None if mask_gt is None else mask_gt.detach().contiguous().float()
        unpacked = torch.empty(B, H, W, 3, dtype=torch.float32, device=x.device)
        sums = torch.empty(2, dtype=torch.float32, device=x.device)
        call("gom_photometric_forward", GomPhotoArgs(
            n_frames=B, height=H, width=W, rgb=ptr(x), rgb_pixel_stride=4, mask=ptr(x[..., 3]), mask_pixel_stride=4,
            bgcolor=ptr(bg), gt_rgb=ptr(gt_rgb), gt_mask=ptr(gt_mask), unpacked=ptr(unpacked), loss_sums=ptr(sums)))
        ctx.save_for_backward(x, bg, gt_rgb, gt_mask)
        n = float(B * H * W)
        return unpacked, sums[0] / (3.0 * n), sums[1] / n

    @staticmethod
    def backward(ctx, g_unpacked, g_lrgb, g_lmask):
        x, bg, gt_rgb, gt_mask = ctx.saved_tensors
        B, H, W, _ = x.shape
        z = torch.zeros((), device=x.device)
        g_loss = torch.stack([z if g_lrgb is None else g_lrgb.float(), z if g_lmask is None else g_lmask.float()])
        g_u = None if g_unpacked is None else g_unpacked.contiguous().float()
        d = torch.empty(B, H, W, 4, dtype=torch.float32, device=x.device)
        call("gom_photometric_backward", GomPhotoArgs(
            n_frames=B, height=H, width=W, rgb=ptr(x), rgb_pixel_stride=4, mask=ptr(x[..., 3]), mask_pixel_stride=4,
            bgcolor=ptr(bg), gt_rgb=ptr(gt_rgb), gt_mask=ptr(gt_mask), dL_dunpacked=ptr(g_u), dL_dlosses=ptr(g_loss),
            dL_drgb=ptr(d), dL_drgb_pixel_stride=4, dL_dmask=ptr(d[..., 3]), dL_dmask_pixel_stride=4))
        return d, None, None, None


class _ShadeRGBA(torch.autograd.Function):
    """(rgba [B,H,W,4] contiguous, shading [B,H,W,1]) -> (rgba[..., :3] * shading, rgba[..., 3]) in one launch each way
    (reference models/model.py:281-287).  The gradient reaches the rasterizer as ONE [B,H,W,4] tensor."""

    @staticmethod
    def forward(ctx, rgba, shading):
        B, H, W, _ = rgba.shape
        x, s = rgba.detach(), shading.detach().reshape(B, H, W).contiguous().float()
        rgbs = torch.empty(B, H, W, 3, dtype=torch.float32, device=x.device)
        masks = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
        call("gom_shade_forward", GomShadeArgs(n_pixels=B * H * W, rgba=ptr(x), shading=ptr(s), rgbs=ptr(rgbs), masks=ptr(masks)))
        ctx.save_for_backward(x, s)
        ctx.shading_shape = tuple(shading.shape)
        return rgbs, masks

    @staticmethod
    def backward(ctx, g_rgbs, g_masks):
        x, s = ctx.saved_tensors
        B, H, W, _ = x.shape
        g_rgbs = None if g_rgbs is None else g_rgbs.contiguous().float()
        g_masks = None if g_masks is None else g_masks.contiguous().float()
        d = torch.empty(B, H, W, 4, dtype=torch.float32, device=x.device)
        ds = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
        call("gom_shade_backward", GomShadeArgs(n_pixels=B * H * W, rgba=ptr(x), shading=ptr(s), dL_drgbs=ptr(g_rgbs),
                                                dL_dmasks=ptr(g_masks), dL_drgba=ptr(d), dL_dshading=ptr(ds)))
        return d, ds.reshape(ctx.shading_shape)


def shade_rgba(rgba, shading):
    """rgba [B,H,W,4] (the rasterizer's output), shading [B,H,W,1] -> (rgbs [B,H,W,3] = albedo * shading, masks [B,H,W])."""
    if not rgba.is_cuda:
        raise _lib.GomError("shade_rgba: inputs must live on a CUDA device (no CPU path exists)")
    if rgba.dtype == torch.float32 and rgba.is_contiguous() and rgba.dim() == 4 and rgba.shape[-1] == 4 \
            and shading.numel() == rgba.numel() // 4:
        return _ShadeRGBA.apply(rgba, shading)
    return rgba[..., :3] * shading, rgba[..., 3]          # unusual layouts / dtypes: the reference's own expression, on the device


def _rgba_base(rgbs, masks):
    """The contiguous fp32 [B,H,W,4] tensor of which ``rgbs`` / ``masks`` are the channel slices [..., :3] / [..., 3]
    (``Model.forward`` without a shadow module returns exactly these views of the rasterizer's output), else None."""
    base = getattr(rgbs, "_base", None)
    if base is None or base is not getattr(masks, "_base", None) or base.dim() != 4 or base.shape[-1] != 4:
        return None
    if base.dtype != torch.float32 or not base.is_cuda or not base.is_contiguous():
        return None
    if tuple(rgbs.shape) != tuple(base.shape[:3]) + (3,) or tuple(masks.shape) != tuple(base.shape[:3]):
        return None
    if rgbs.stride() != base.stride() or rgbs.storage_offset() != base.storage_offset():
        return None
    if masks.stride() != base.stride()[:3] or masks.storage_offset() != base.storage_offset() + 3:
        return None
    return base


def photometric_l1(rgbs, masks, bgcolors, rgb_gt, mask_gt):
    """rgbs [B,H,W,3], masks [B,H,W], bgcolors [B,3] or None, rgb_gt [B,H,W,3], mask_gt [B,H,W]  ->
    (rgb_unpacked [B,H,W,3], mean|rgb_unpacked - rgb_gt|, mean|masks - mask_gt|)   (train.py:53-55,101-111)"""
    base = _rgba_base(rgbs, masks)
    if base is not None:
        return _PhotometricL1RGBA.apply(base, bgcolors, rgb_gt, mask_gt)
    return _PhotometricL1.apply(rgbs, masks, bgcolors, rgb_gt, mask_gt)


def unpack(rgbs, masks, bgcolors):
    """reference train.py:53-55: rgbs * masks + bgcolors * (1 - masks)."""
    return _PhotometricL1.apply(rgbs, masks, bgcolors, None, None)[0]


def compute_loss(rgbs, masks, bgcolors, rgb_gt, mask_gt, lpips_func=None, coeff_rgb=1.0, coeff_mask=5.0, coeff_lpips=1.0):
    """The photometric part of reference train.py:98-121 (with `unpack` of :325-326 folded in).  Returns
    (total, {'rgb','mask','lpips'} unscaled terms, rgb_unpacked)."""
    rgb_u, l_rgb, l_mask = photometric_l1(rgbs, masks, bgcolors, rgb_gt, mask_gt)
    total = coeff_rgb * l_rgb + coeff_mask * l_mask
    terms = {"rgb": l_rgb, "mask": l_mask}
    if lpips_func is not None and coeff_lpips > 0:
        if getattr(lpips_func, "fused", False) and hasattr(lpips_func, "per_image"):
            l_lp = torch.mean(lpips_func.per_image(rgb_u, rgb_gt, from_unit_range=True))       # 2x-1 folded into the kernel
        else:
            s = lambda x: 2 * x - 1
            l_lp = torch.mean(lpips_func(s(rgb_u.permute(0, 3, 1, 2)), s(rgb_gt.permute(0, 3, 1, 2))))
        terms["lpips"] = l_lp
        total = total + coeff_lpips * l_lp
    return total, terms, rgb_u
