"""3x3 convolutions (stride 1, zero padding 1) over NHWC fp32 activations on the tcgen05 implicit-GEMM kernel of
csrc/conv3x3_tc.cu — the VGG16 layers conv1_2 ... conv5_3 of LPIPS (reference utils/lpips/pretrained_networks.py:96-134),
forward with bias + ReLU fused and input gradient with the ReLU backward of the layer below fused.

There is no CPU / cuDNN path here: every function raises ``GomError`` when the library is missing or a tensor is not on a
CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomConv3x3Args, GomConvPackArgs, GomTf32SplitArgs, call, ptr


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise _lib.GomError(f"{what}: tensors must live on a CUDA device (no CPU path exists)")


def pack_weights(weight, transpose=False, split=False):
    """torch weight [K,C,3,3] -> the kernel's packed image: [9,K,C] (forward) or [9,C,K] with flipped taps (dgrad),
    TF32-rounded; ``split`` appends the low parts (3xTF32)."""
    _need_cuda(weight, "pack_weights")
    K, C = weight.shape[:2]
    w = weight.detach().float().contiguous()
    packed = torch.empty((18 if split else 9), (C if transpose else K), (K if transpose else C), dtype=torch.float32,
                         device=weight.device)
    call("gom_conv3x3_pack_weights", GomConvPackArgs(c_out=K, c_in=C, transpose=int(transpose), split=int(split),
                                                     weight=ptr(w), packed=ptr(packed)))
    return packed


def tf32_low_part(x):
    """x - trunc_tf32(x): what the tensor core drops when it reads x as a TF32 operand (second operand of 3xTF32)."""
    _need_cuda(x, "tf32_low_part")
    lo = torch.empty_like(x)
    call("gom_tf32_split", GomTf32SplitArgs(n=x.numel(), x=ptr(x), hi=None, lo=ptr(lo)))
    return lo


def conv3x3(x, w_packed, bias=None, relu=False, mask_in=None, mask_out=None, precision="tf32", tma_round=True, out=None,
            status=None):
    """x: contiguous [N,H,W,C_in]; w_packed from ``pack_weights``; returns [N,H,W,C_out] = conv (+ bias) (ReLU) (masked).

    ``mask_out`` (int32 [N,H,W,C_out/32]) receives the ReLU bit mask of the result; ``mask_in`` (same shape) zeroes the
    result where its bits are 0 (the fused ReLU backward of the dgrad)."""
    _need_cuda(x, "conv3x3")
    if not x.is_contiguous():
        raise _lib.GomError("conv3x3: x must be a contiguous NHWC tensor")
    N, H, W, C = x.shape
    c_out = w_packed.shape[1]
    if w_packed.shape[2] != C:
        raise _lib.GomError(f"conv3x3: packed weight expects {w_packed.shape[2]} input channels, x has {C}")
    strict = precision in ("fp32", "3xtf32")
    if strict and w_packed.shape[0] != 18:
        raise _lib.GomError("conv3x3: 3xTF32 needs a weight packed with split=True")
    if out is None:
        out = torch.empty(N, H, W, c_out, dtype=torch.float32, device=x.device)
    x_lo = tf32_low_part(x) if strict else None
    for mk in (mask_in, mask_out):
        if mk is not None and (tuple(mk.shape) != (N, H, W, c_out // 32) or mk.dtype != torch.int32 or not mk.is_contiguous()):
            raise _lib.GomError("conv3x3: masks must be contiguous int32 [N,H,W,C_out/32]")
    call("gom_conv3x3", GomConv3x3Args(n_images=N, height=H, width=W, c_in=C, c_out=c_out, relu=int(relu), precision=int(strict),
                                       tma_round=int(tma_round and not strict), x=ptr(x), x_lo=ptr(x_lo), w_packed=ptr(w_packed),
                                       bias=ptr(bias), mask_in=ptr(mask_in), mask_out=ptr(mask_out), out=ptr(out), status=ptr(status)))
    return out


def new_mask(n, h, w, c, device):
    """storage for the ReLU bit mask of an [n,h,w,c] activation"""
    return torch.empty(n, h, w, c // 32, dtype=torch.int32, device=device)
