"""LPIPS-VGG v0.1 perceptual loss (reference utils/lpips/lpips.py:81-123, pretrained_networks.py:96-134,
__init__.py:40-42; used at train.py:113-121 and eval.py:110-116).

Every operation of the loss runs in hand-written kernels of libgom_b200.so (``fused=True``, the default):
  * the VGG16 3x3 convolutions: conv1_1 in csrc/conv_first_tc.cu, conv1_2 ... conv5_3 in csrc/conv3x3_tc.cu (tcgen05
    implicit GEMMs fed by TMA, bias + ReLU fused into the forward epilogue, the ReLU backward of the layer below fused
    into the dgrad epilogue through a bit mask the forward wrote);
  * input scaling, max-pooling, channel normalisation, squared difference, the 1x1 heads, the spatial mean and all of
    their backward passes in csrc/lpips.cu (one HBM pass per tapped layer each way instead of ~25),
driven by the hand-rolled backward in ``_FusedLpips``: prediction and target go through the trunk as ONE batch
[B pred | B gt], only the prediction half is back-propagated.  ``conv_impl="cudnn"`` keeps the library convolutions
(the A/B baseline of tools/conv_probe.py and of the tests; not the product path); ``fused=False`` keeps the plain
torch formulation (the A/B reference for the kernels' tests).  Numerics switch ``conv_precision``:
  "tf32" (default) — TF32 tensor-core products, fp32 accumulation.  This IS the reference's stock behaviour: it never
          touches ``torch.backends.cudnn.allow_tf32``, whose default is True in the torch 1.13 it pins (README.md:19-20);
          operands are rounded to TF32 to nearest (weights when packed, activations by the TMA engine), like cuDNN's;
  "fp32"  — 3xTF32 products (fp32-GEMM accuracy; what the CPU oracle computes; used by the parity tests);
  "bf16"  — autocast to bfloat16 on the torch path (below the 1e-3 gradient tolerance: opt-in only).

Weights: the trunk is torchvision's VGG16 ``features[:30]`` (same state-dict keys); ImageNet weights are not
downloadable offline, so ``trunk_state`` must be given (or ``seeded_random_trunk`` used for tests/benchmarks, like the
reference's ``pnet_rand=True``).  ``head_weights`` are the five 1x1 heads of ``utils/lpips/weights/v0.1/vgg.pth``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import conv as _conv
from ._lib import GomBiasReluArgs, GomConvFirstArgs, GomLpipsInputArgs, GomLpipsTapArgs, GomReluBwdArgs, call, ptr

_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_TAPS = (3, 8, 15, 22, 29)
CHANNELS = (64, 128, 256, 512, 512)


def make_vgg16_features():
    layers, cin = [], 3
    for v in _VGG_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=False)]
            cin = v
    return nn.Sequential(*layers)


def seeded_random_trunk(seed=0):
    """State of the trunk that the reference's ``LPIPS(net='vgg', pnet_rand=True)`` builds right after
    ``torch.manual_seed(seed)`` (torchvision initialises the full VGG16, features first)."""
    import torchvision
    torch.manual_seed(seed)
    net = torchvision.models.vgg16(weights=None)
    return {k: v.clone() for k, v in net.features.state_dict().items() if int(k.split(".")[0]) < 30}


def load_head_weights(path):
    """Read the five ``linK.model.1.weight`` tensors of the reference's ``utils/lpips/weights/v0.1/vgg.pth``."""
    sd = torch.load(path, map_location="cpu")
    return [sd[f"lin{k}.model.1.weight"].reshape(-1) for k in range(5)]


class _BackwardPrecision:
    """Makes the cuDNN TF32 switch that was active in the forward also govern the trunk's backward convolutions (they run
    later, inside loss.backward()): an identity on every tap output flips the flag when the backward pass reaches the
    trunk, an identity on the trunk input restores it once the backward has left the trunk."""

    class _Enter(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, state, allow):
            ctx.state, ctx.allow = state, allow
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            if "prev" not in ctx.state:
                ctx.state["prev"] = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = ctx.allow
            return g, None, None

    class _Exit(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, state):
            ctx.state = state
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            if "prev" in ctx.state:
                torch.backends.cudnn.allow_tf32 = ctx.state.pop("prev")
            return g, None


_LEVELS = (2, 2, 3, 3, 3)           # convolutions per VGG16 block; the last ReLU of each block is tapped


def _nhwc(t):
    """[N,C,H,W] channels_last tensor -> the same storage seen as contiguous [N,H,W,C]"""
    return t.permute(0, 2, 3, 1)


class _FusedLpips(torch.autograd.Function):
    """pred, gt: contiguous [B,H,W,3] fp32 -> per-image LPIPS value [B].  Only ``pred`` receives a gradient.
    Activations are contiguous NHWC tensors [2B,h,w,C] throughout."""

    @staticmethod
    def forward(ctx, pred, gt, net, from_unit_range):
        if pred.device.type != "cuda":
            raise _lib.GomError("LPIPS (fused): inputs must live on a CUDA device (no CPU path exists)")
        B, H, W, _ = pred.shape
        dev = pred.device
        pred, gt = pred.detach().contiguous().float(), gt.detach().contiguous().float()
        x = torch.empty(2 * B, H, W, 3, dtype=torch.float32, device=dev)
        call("gom_lpips_input_forward", GomLpipsInputArgs(n_frames=B, height=H, width=W, from_unit_range=int(from_unit_range),
                                                          pred=ptr(pred), gt=ptr(gt), out=ptr(x)))
        vals = torch.zeros(B, dtype=torch.float32, device=dev)
        acts, masks = [x], {}
        h = x
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = net.conv_precision != "fp32"
        try:
            ci = 0
            for level, n_conv in enumerate(_LEVELS):
                for j in range(n_conv):
                    # the ReLU mask of an activation that feeds another convolution of the same block is what that
                    # convolution's dgrad needs (the block's last activation is tapped: csrc/lpips.cu masks there)
                    h, m = net._conv_bias_relu(h, ci, want_mask=j < n_conv - 1)
                    if m is not None:
                        masks[len(acts)] = m
                    acts.append(h)
                    ci += 1
                N2, hh, ww, C = h.shape
                pool = level < len(_LEVELS) - 1
                pooled = torch.empty(N2, hh // 2, ww // 2, C, dtype=torch.float32, device=dev) if pool else None
                call("gom_lpips_tap_forward", GomLpipsTapArgs(
                    n_frames=B, height=hh, width=ww, channels=C, pool=int(pool), feats=ptr(h), lin=ptr(net._lin(level)),
                    layer_sums=ptr(vals), pooled=ptr(pooled)))
                if pool:
                    h = pooled
                    acts.append(h)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        ctx.net, ctx.acts, ctx.masks, ctx.from_unit_range = net, acts, masks, bool(from_unit_range)
        ctx.dims = (B, H, W)
        return vals

    @staticmethod
    def backward(ctx, g_vals):
        net, acts, masks = ctx.net, ctx.acts, ctx.masks
        B, H, W = ctx.dims
        dev = g_vals.device
        dval = g_vals.contiguous().float()
        # index of the activation feeding / produced by every convolution, walking acts = [x, c0, c1, pool, c2, ...]
        conv_in, conv_out, k = [], [], 0
        for level, n_conv in enumerate(_LEVELS):
            for _ in range(n_conv):
                conv_in.append(k); conv_out.append(k + 1); k += 1
            if level < len(_LEVELS) - 1:
                k += 1
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = net.conv_precision != "fp32"
        try:
            g_pooled, ci, fuse_act = None, len(conv_in), None
            for level in reversed(range(len(_LEVELS))):
                ci_last = ci - 1
                a = acts[conv_out[ci_last]]
                N2, hh, ww, C = a.shape
                g_pre = torch.empty(B, hh, ww, C, dtype=torch.float32, device=dev)
                call("gom_lpips_tap_backward", GomLpipsTapArgs(
                    n_frames=B, height=hh, width=ww, channels=C, pool=int(g_pooled is not None), feats=ptr(a),
                    lin=ptr(net._lin(level)), dL_dval=ptr(dval), dL_dpooled=ptr(g_pooled), dL_dpre=ptr(g_pre)))
                for j in reversed(range(_LEVELS[level])):
                    ci -= 1
                    inp = acts[conv_in[ci]][:B]
                    mask = masks.get(conv_in[ci]) if j > 0 else None       # fused ReLU backward of the layer below
                    g_in = net._conv_dgrad(g_pre, inp, ci, act=fuse_act, mask_in=None if mask is None else mask[:B])
                    fuse_act = None
                    if j > 0:                                # the input was itself a ReLU output of this block
                        if mask is None:
                            if ci == 1 and net.own_first_conv and net.first_conv_tc:
                                # conv1_1's output: its ReLU backward is fused into the tcgen05 dgrad that consumes this
                                # gradient next (one read + one write of the largest gradient tensor and a launch less)
                                fuse_act = inp
                            else:
                                call("gom_relu_backward", GomReluBwdArgs(n=g_in.numel(), act=ptr(inp), grad=ptr(g_in)))
                        g_pre = g_in
                    else:                                    # the input was the pooled previous block (or the image)
                        g_pooled = g_in
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        d_pred = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev)
        call("gom_lpips_input_backward", GomLpipsInputArgs(n_frames=B, height=H, width=W, from_unit_range=int(ctx.from_unit_range),
                                                           dL_dout=ptr(g_pooled), dL_dpred=ptr(d_pred)))
        ctx.acts = ctx.masks = None
        return d_pred, None, None, None


class LPIPS(nn.Module):
    def __init__(self, trunk_state, head_weights, conv_precision="tf32", channels_last=True, fused=True,
                 conv_epilogue="cudnn", conv_impl="tcgen05", streams=1):
        super().__init__()
        # streams > 1: the batch is cut into that many groups of frames, each taken through the whole network (forward and,
        # through autograd, backward) on its own CUDA stream.  The step alternates tensor-bound convolutions (one persistent CTA
        # per SM, ~200 KB of shared memory, 36 k registers) with HBM-bound tap / input / first-layer kernels (80 registers, no
        # shared memory): with two groups in flight a group's HBM-bound kernel runs next to the other group's convolution on
        # the same SMs instead of after it.  Per-image results do not depend on the grouping.
        self.streams = int(streams)
        self._side_streams = {}
        self.fused = bool(fused) and conv_precision != "bf16"      # the fused kernels are fp32-only
        assert conv_epilogue in ("kernel", "cudnn")
        assert conv_impl in ("tcgen05", "cudnn")
        self.conv_epilogue = conv_epilogue       # cuDNN baseline only: its fused bias + ReLU epilogue, or csrc/lpips.cu's
        self.conv_impl = conv_impl
        self._packs = {}
        self.features = make_vgg16_features()
        self.features.load_state_dict(trunk_state)
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])
        for k, w in enumerate(head_weights):
            self.register_buffer(f"lin{k}", torch.as_tensor(np.asarray(w), dtype=torch.float32).reshape(1, -1, 1, 1))
        assert conv_precision in ("tf32", "fp32", "bf16")
        self.conv_precision = conv_precision
        self.channels_last = channels_last
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
        if channels_last:
            self.features = self.features.to(memory_format=torch.channels_last)
        self._convs = [m for m in self.features if isinstance(m, nn.Conv2d)]
        self._no_cudnn_epilogue = set()
        # conv1_1 (3 -> 64) runs in csrc/conv_first.cu: torch-contiguous [64,3,3,3] copy of its weight
        self.own_first_conv = True
        self.first_conv_tc = True      # csrc/conv_first_tc.cu (tcgen05, HBM-bound) instead of csrc/conv_first.cu (FFMA-bound)
        self.register_buffer("w_first", self._convs[0].weight.detach().clone(memory_format=torch.contiguous_format),
                             persistent=False)

    # ---------------------------------------------------------------------------- fused path (csrc/lpips.cu + cuDNN)
    def _lin(self, level):
        return getattr(self, f"lin{level}").reshape(-1)

    def _packed(self, ci, transpose):
        """packed weight image of convolution ``ci`` for csrc/conv3x3_tc.cu (built once per device and precision)"""
        w = self._convs[ci].weight
        strict = self.conv_precision == "fp32"
        key = (ci, transpose, strict, w.device)
        if key not in self._packs:
            self._packs[key] = _conv.pack_weights(w, transpose=transpose, split=strict)
        return self._packs[key]

    def _conv_bias_relu(self, h, ci, want_mask=False):
        """3x3 convolution + bias + ReLU on a contiguous NHWC batch -> (activation, ReLU bit mask or None)."""
        conv = self._convs[ci]
        N, hh, ww, _ = h.shape
        if ci == 0 and self.own_first_conv:
            y = torch.empty(N, hh, ww, 64, dtype=torch.float32, device=h.device)
            # the ReLU bit mask of conv1_1's output, in the layout conv1_2's dgrad applies (the backward of conv1_1 then reads
            # neither its own 1 GB activation nor a mask)
            mask = _conv.new_mask(N, hh, ww, 64, h.device) if (want_mask and self.first_conv_tc and self.conv_impl == "tcgen05") else None
            call("gom_conv_first_forward", GomConvFirstArgs(n_images=N, height=hh, width=ww, use_tensor_cores=int(self.first_conv_tc),
                                                            x=ptr(h), weight=ptr(self.w_first), bias=ptr(conv.bias), out=ptr(y),
                                                            mask_out=ptr(mask)))
            return y, mask
        if self.conv_impl == "tcgen05" and ci > 0:
            mask = _conv.new_mask(N, hh, ww, conv.out_channels, h.device) if want_mask else None
            y = _conv.conv3x3(h, self._packed(ci, False), bias=conv.bias, relu=True, mask_out=mask, precision=self.conv_precision)
            return y, mask
        # cuDNN baseline (channels_last views of the same NHWC storage)
        w = conv.weight
        hn = h.permute(0, 3, 1, 2)
        if self.conv_epilogue == "cudnn" and ci not in self._no_cudnn_epilogue:
            try:
                y = torch.cudnn_convolution_relu(hn, w, conv.bias, (1, 1), (1, 1), (1, 1), 1)
                return y.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1), None
            except RuntimeError:
                self._no_cudnn_epilogue.add(ci)          # unsupported shape: use the kernel epilogue for this layer
        y = F.conv2d(hn, w, None, padding=1)
        y = y.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
        call("gom_bias_relu", GomBiasReluArgs(n_pixels=N * hh * ww, channels=conv.out_channels, x=ptr(y), bias=ptr(conv.bias)))
        return y, None

    def _conv_dgrad(self, g_out, inp, ci, act=None, mask_in=None):
        """Input gradient of convolution ``ci`` (NHWC in, NHWC out).  ``act``: (first convolution, tensor-core kernel only) its
        own ReLU output; g_out is then the unmasked gradient.  ``mask_in``: ReLU bit mask of ``inp`` (tcgen05 path)."""
        N, hh, ww, _ = g_out.shape
        if ci == 0 and self.own_first_conv:
            dx = torch.empty(N, hh, ww, 3, dtype=torch.float32, device=g_out.device)
            scratch = torch.empty(9, N * hh * ww, 4, dtype=torch.float32, device=g_out.device) if self.first_conv_tc else None
            call("gom_conv_first_backward", GomConvFirstArgs(n_images=N, height=hh, width=ww, use_tensor_cores=int(self.first_conv_tc),
                                                             weight=ptr(self.w_first), dL_dout=ptr(g_out), dL_dx=ptr(dx), scratch=ptr(scratch),
                                                             act=ptr(act)))
            return dx
        if self.conv_impl == "tcgen05" and ci > 0:
            return _conv.conv3x3(g_out, self._packed(ci, True), mask_in=mask_in, precision=self.conv_precision)
        w = self._convs[ci].weight
        g = torch.ops.aten.convolution_backward(g_out.permute(0, 3, 1, 2), inp.permute(0, 3, 1, 2), w, None, (1, 1), (1, 1), (1, 1),
                                                False, (0, 0), 1, (True, False, False))[0]
        return g.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)

    def per_image(self, pred_nhwc, gt_nhwc, from_unit_range=True):
        """pred / gt contiguous [B,H,W,3] (in [0,1] when from_unit_range, else already in [-1,1]) -> LPIPS values [B]."""
        B = pred_nhwc.shape[0]
        n = min(self.streams, B)
        if n <= 1 or not pred_nhwc.is_cuda:
            return _FusedLpips.apply(pred_nhwc, gt_nhwc, self, from_unit_range)
        dev = pred_nhwc.device
        if dev not in self._side_streams or len(self._side_streams[dev]) < n:
            self._side_streams[dev] = [torch.cuda.Stream(device=dev) for _ in range(n)]
        cur = torch.cuda.current_stream(dev)
        bounds = [B * i // n for i in range(n + 1)]
        outs = []
        for i in range(n):
            st = self._side_streams[dev][i]
            st.wait_stream(cur)                                  # fork (inside a CUDA-graph capture: joins the capture)
            with torch.cuda.stream(st):
                p, g = pred_nhwc[bounds[i]:bounds[i + 1]], gt_nhwc[bounds[i]:bounds[i + 1]]
                o = _FusedLpips.apply(p, g, self, from_unit_range)
                o.record_stream(cur)
                outs.append(o)
        for i in range(n):
            cur.wait_stream(self._side_streams[dev][i])          # join
        return torch.cat(outs)

    def _taps(self, x):
        h = (x - self.shift) / self.scale
        if self.channels_last:
            h = h.contiguous(memory_format=torch.channels_last)
        outs = []
        for i, layer in enumerate(self.features):
            h = layer(h)
            if i in _TAPS:
                outs.append(h)
        return outs

    @staticmethod
    def _unit(f, eps=1e-10):
        n = torch.sqrt(torch.sum(f * f, dim=1, keepdim=True) + eps)
        return f / (n + eps)

    def target_features(self, in1):
        """Unit-normalised features of the (gradient-free) target image; reusable across calls."""
        with torch.no_grad():
            return [self._unit(f) for f in self._run(self._taps, in1)]

    def _run(self, fn, *a):
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = self.conv_precision != "fp32"
        try:
            if self.conv_precision == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return [o.float() for o in fn(*a)]
            return fn(*a)
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def forward(self, in0, in1=None, target_feats=None):
        """in0 / in1 in [-1,1], [B,3,H,W] -> [B,1,1,1]  (reference LPIPS.forward with normalize=False)."""
        if self.fused and target_feats is None and self.conv_precision != "bf16":
            B = in0.shape[0]
            return self.per_image(in0.permute(0, 2, 3, 1), in1.permute(0, 2, 3, 1), from_unit_range=False).view(B, 1, 1, 1)
        if target_feats is None:
            target_feats = self.target_features(in1)
        if in0.requires_grad:
            state = {}
            allow = self.conv_precision != "fp32"
            in0 = _BackwardPrecision._Exit.apply(in0, state)
            feats0 = [_BackwardPrecision._Enter.apply(f, state, allow) for f in self._run(self._taps, in0)]
        else:
            feats0 = self._run(self._taps, in0)
        total = 0
        for k, (f0, f1) in enumerate(zip(feats0, target_feats)):
            d = (self._unit(f0) - f1) ** 2
            total = total + (d * getattr(self, f"lin{k}")).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
        return total
