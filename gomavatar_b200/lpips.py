"""LPIPS-VGG v0.1 perceptual loss (reference utils/lpips/lpips.py:81-123, pretrained_networks.py:96-134,
__init__.py:40-42; used at train.py:113-121 and eval.py:110-116).

The VGG16 convolutions are dense GEMM-shaped work and stay in cuDNN (library tensor-core kernels, as SURVEY.md §8a-12
prescribes); what this module owns is the glue: layout (channels_last), the target-image feature cache (the ground
truth of a frame does not change between the forward and anything else in the step, and carries no gradient), and
the numerics switch ``conv_precision``:
  "tf32" (default) — cuDNN may use TF32 tensor-core convolutions.  This IS the reference's stock behaviour: it never
          touches ``torch.backends.cudnn.allow_tf32``, whose default is True in the torch 1.13 it pins (README.md:19-20);
  "fp32"  — strict IEEE fp32 convolutions (what the CPU oracle computes; used by the parity tests);
  "bf16"  — autocast to bfloat16 (fastest, below the 1e-3 gradient tolerance: opt-in only).

Weights: the trunk is torchvision's VGG16 ``features[:30]`` (same state-dict keys); ImageNet weights are not
downloadable offline, so ``trunk_state`` must be given (or ``seeded_random_trunk`` used for tests/benchmarks, like the
reference's ``pnet_rand=True``).  ``head_weights`` are the five 1x1 heads of ``utils/lpips/weights/v0.1/vgg.pth``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

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


class LPIPS(nn.Module):
    def __init__(self, trunk_state, head_weights, conv_precision="tf32", channels_last=True):
        super().__init__()
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
