"""Oracle (test infrastructure): photometric losses and eval metrics of the reference, restated on torch/numpy CPU.

* ``unpack``                reference train.py:53-55 (eval variant with clamp: eval.py:80-83)
* ``l1_losses``             reference train.py:101-111
* ``LPIPSVGG``              reference utils/lpips/lpips.py:81-123,126-146, pretrained_networks.py:96-134,
                            __init__.py:40-42  (v0.1, net='vgg', spatial=False, eval mode)
* ``psnr`` / ``ssim``       reference eval.py:101-108 with skimage 0.18 defaults (requirements.txt:12): 7x7 uniform
                            window, sample covariance, K1=.01 K2=.03, data_range=2 for float input, 3-px crop.
                            skimage is absent offline: PARITY UNPINNED for ssim (restated from published defaults).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def unpack(rgbs, masks, bgcolors, clamp=False):
    out = rgbs * masks.unsqueeze(-1) + bgcolors[:, None, None, :] * (1 - masks).unsqueeze(-1)
    return out.clamp(0.0, 1.0) if clamp else out


def l1_losses(rgb_pred, mask_pred, rgb_gt, mask_gt):
    return torch.mean(torch.abs(rgb_pred - rgb_gt)), torch.mean(torch.abs(mask_pred - mask_gt))


_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_TAPS = (3, 8, 15, 22, 29)          # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 in torchvision's numbering
_CHNS = (64, 128, 256, 512, 512)


def make_vgg16_features():
    """torchvision ``vgg16().features[:30]`` topology, built without torchvision (indices match its state_dict)."""
    layers, cin = [], 3
    for v in _VGG_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=False)]
            cin = v
    return nn.Sequential(*layers)


def seeded_random_trunk_state(seed=0):
    """The trunk the reference gets from ``LPIPS(net='vgg', pnet_rand=True)`` right after
    ``torch.manual_seed(seed)`` (torchvision initialises the whole VGG16, features first)."""
    import torchvision
    torch.manual_seed(seed)
    net = torchvision.models.vgg16(weights=None)
    return {k: v.clone() for k, v in net.features.state_dict().items() if int(k.split(".")[0]) < 30}


class LPIPSVGG(nn.Module):
    def __init__(self, trunk_state, head_weights):
        super().__init__()
        self.features = make_vgg16_features()
        self.features.load_state_dict(trunk_state)
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])
        self.heads = [torch.as_tensor(np.asarray(w), dtype=torch.float32).reshape(1, -1, 1, 1) for w in head_weights]
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def _taps(self, x):
        outs, h = [], (x - self.shift) / self.scale
        for i, layer in enumerate(self.features):
            h = layer(h)
            if i in _TAPS:
                outs.append(h)
        return outs

    @staticmethod
    def _unit(f, eps=1e-10):
        n = torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True) + eps)
        return f / (n + eps)

    def forward(self, in0, in1):
        """inputs in [-1,1], [B,3,H,W] -> [B,1,1,1]"""
        total = 0
        for f0, f1, w in zip(self._taps(in0), self._taps(in1), self.heads):
            d = (self._unit(f0) - self._unit(f1)) ** 2
            total = total + (d * w.to(d.dtype)).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
        return total


def lpips_loss(lpips_mod, rgb_pred, rgb_gt):
    """reference train.py:113-117: inputs [B,H,W,3] in [0,1]."""
    s = lambda x: 2 * x - 1
    return torch.mean(lpips_mod(s(rgb_pred.permute(0, 3, 1, 2)), s(rgb_gt.permute(0, 3, 1, 2))))


def to_8b(img):
    """reference utils/image_util.py:21-22."""
    return (255.0 * np.clip(img, 0.0, 1.0)).astype(np.uint8)


def psnr(pred, gt):
    """reference eval.py:101-104 on float images in [0,1]."""
    mse = np.mean((pred - gt) ** 2)
    return -10.0 * np.log(mse) / np.log(10.0)


def ssim(pred, gt, win=7, data_range=2.0, K1=0.01, K2=0.03):
    """skimage 0.18 ``structural_similarity(pred, gt, multichannel=True)`` defaults on float64 [H,W,C] images."""
    pred = np.asarray(pred, np.float64)
    gt = np.asarray(gt, np.float64)
    H, W, C = pred.shape
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    pad = (win - 1) // 2

    def box(a):   # uniform filter, evaluated only where the window fits (the crop discards the rest)
        cs = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), axis=0), axis=1)
        s = cs[win:, win:] - cs[:-win, win:] - cs[win:, :-win] + cs[:-win, :-win]
        return s / NP

    vals = []
    for ch in range(C):
        x, y = pred[..., ch], gt[..., ch]
        ux, uy = box(x), box(y)
        vx = cov_norm * (box(x * x) - ux * ux)
        vy = cov_norm * (box(y * y) - uy * uy)
        vxy = cov_norm * (box(x * y) - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        assert S.shape == (H - 2 * pad, W - 2 * pad)
        vals.append(S.mean())
    return float(np.mean(vals))
