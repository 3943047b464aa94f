"""3x3 convolutions (stride 1, zero padding 1) over NHWC fp32 activations on the tcgen05 implicit-GEMM kernel of
csrc/conv3x3_tc.cu — the VGG16 layers conv1_2 ... conv5_3 of LPIPS (reference utils/lpips/pretrained_networks.py:96-134),
forward with bias + ReLU fused and input gradient with the ReLU backward of the layer below fused.

There is no CPU / cuDNN path here: every function raises ``GomError`` when the library is missing or a tensor is not on a
CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomConv3x3Args, GomConvPackArgs, GomLinearWgradArgs, GomTf32SplitArgs, call, ptr


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise _lib.GomError(f"{what}: tensors must live on a CUDA device (no CPU path exists)")


def pack_weights(weight, transpose=False, split=False):
    """torch weight [K,C,3,3] (or [K,C]: a Linear layer) -> the kernel's packed image: [taps,K,C] (forward) or [taps,C,K]
    with flipped taps (dgrad), TF32-rounded; ``split`` appends the low parts (3xTF32)."""
    _need_cuda(weight, "pack_weights")
    K, C = weight.shape[:2]
    ks = 1 if weight.dim() == 2 else int(weight.shape[2])
    taps = ks * ks
    w = weight.detach().float().contiguous()
    packed = torch.empty((2 * taps if split else taps), (C if transpose else K), (K if transpose else C), dtype=torch.float32,
                         device=weight.device)
    call("gom_conv3x3_pack_weights", GomConvPackArgs(c_out=K, c_in=C, transpose=int(transpose), split=int(split), kernel_size=ks,
                                                     weight=ptr(w), packed=ptr(packed)))
    return packed


def tf32_low_part(x, col_sum=None):
    """x - trunc_tf32(x): what the tensor core drops when it reads x as a TF32 operand (second operand of 3xTF32).
    ``col_sum`` (fp32 [x.shape[-1]], a multiple of 4 dividing 1024) is incremented by the column sums of x in the same pass."""
    _need_cuda(x, "tf32_low_part")
    if not x.is_contiguous():
        raise _lib.GomError("tf32_low_part: x must be contiguous")
    lo = torch.empty_like(x)
    call("gom_tf32_split", GomTf32SplitArgs(n=x.numel(), x=ptr(x), hi=None, lo=ptr(lo), col_sum=ptr(col_sum),
                                            n_cols=0 if col_sum is None else int(x.shape[-1])))
    return lo


def conv3x3(x, w_packed, bias=None, relu=False, mask_in=None, mask_out=None, precision="tf32", tma_round=True, out=None,
            status=None, kernel_size=3, x_lo=None):
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
    if strict and w_packed.shape[0] != 2 * kernel_size * kernel_size:
        raise _lib.GomError("conv3x3: 3xTF32 needs a weight packed with split=True")
    if out is None:
        out = torch.empty(N, H, W, c_out, dtype=torch.float32, device=x.device)
    if strict and x_lo is None:
        x_lo = tf32_low_part(x)
    for mk in (mask_in, mask_out):
        if mk is not None and (tuple(mk.shape) != (N, H, W, c_out // 32) or mk.dtype != torch.int32 or not mk.is_contiguous()):
            raise _lib.GomError("conv3x3: masks must be contiguous int32 [N,H,W,C_out/32]")
    call("gom_conv3x3", GomConv3x3Args(n_images=N, height=H, width=W, c_in=C, c_out=c_out, relu=int(relu), precision=int(strict),
                                       tma_round=int(tma_round and not strict), kernel_size=kernel_size, x=ptr(x), x_lo=ptr(x_lo), w_packed=ptr(w_packed),
                                       bias=ptr(bias), mask_in=ptr(mask_in), mask_out=ptr(mask_out), out=ptr(out), status=ptr(status)))
    return out


def new_mask(n, h, w, c, device):
    """storage for the ReLU bit mask of an [n,h,w,c] activation"""
    return torch.empty(n, h, w, c // 32, dtype=torch.int32, device=device)


def linear(x, w_packed, bias=None, relu=False, mask_in=None, mask_out=None, precision="fp32", out=None, x_lo=None):
    """Linear layer over rows on the tensor cores: x [R, C_in] (R a multiple of 16, C_in a multiple of 32, contiguous) ->
    [R, C_out] = x W^T (+ bias) (ReLU) (masked), C_out a multiple of 64; ``w_packed = pack_weights(W [C_out, C_in], ...)``.
    The rows are handed to the convolution kernel as a 16-pixel-wide one-tap "image" (kernel_size 1).  Masks: int32
    [R, C_out / 32] as in ``conv3x3``."""
    R, C = x.shape
    if R % 16 or C % 32:
        raise _lib.GomError("linear: rows must be a multiple of 16 and input features a multiple of 32 (pad with zeros)")
    c_out = w_packed.shape[1]
    v4 = lambda t, c: None if t is None else t.view(1, R // 16, 16, c)
    if out is None:
        out = torch.empty(R, c_out, dtype=torch.float32, device=x.device)
    conv3x3(x.view(1, R // 16, 16, C), w_packed, bias=bias, relu=relu, mask_in=v4(mask_in, c_out // 32), mask_out=v4(mask_out, c_out // 32),
            precision=precision, out=out.view(1, R // 16, 16, c_out), kernel_size=1,
            x_lo=None if x_lo is None else x_lo.view(1, R // 16, 16, C))
    return out


def linear_wgrad(g, g_lo, x, x_lo, out=None, accumulate=False, status=None):
    """Weight gradient of ``linear``: g [R, 128] (output gradient), x [R, C_in] (the layer's input, C_in a multiple of 32, at most
    256), their TF32 low parts from ``tf32_low_part``  ->  [128, C_in] = g^T x with 3xTF32 products (csrc/wgrad_tc.cu)."""
    _need_cuda(g, "linear_wgrad")
    R, M = g.shape
    N = x.shape[1]
    for t in (g, g_lo, x, x_lo):
        if not t.is_contiguous() or t.dtype != torch.float32 or t.shape[0] != R:
            raise _lib.GomError("linear_wgrad: operands must be contiguous fp32 matrices with the same number of rows")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=g.device)
        accumulate = False
    call("gom_linear_wgrad", GomLinearWgradArgs(rows=R, m=M, n=N, zero_first=int(not accumulate), g=ptr(g), g_lo=ptr(g_lo), x=ptr(x),
                                               x_lo=ptr(x_lo), out=ptr(out), status=ptr(status)))
    return out
