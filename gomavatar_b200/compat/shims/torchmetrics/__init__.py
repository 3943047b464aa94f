"""Stand-in for ``torchmetrics`` (eval.py:24-25, used by ``Evaluator_snapshot`` only; absent offline, PARITY UNPINNED):
``PeakSignalNoiseRatio(data_range)`` and ``StructuralSimilarityIndexMeasure(data_range)`` with torchmetrics' defaults
(11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03, reflect padding cropped again, mean over the map)."""
import torch
import torch.nn.functional as F


class PeakSignalNoiseRatio(torch.nn.Module):
    def __init__(self, data_range=None, **kwargs):
        super().__init__()
        self.data_range = data_range

    def forward(self, preds, target):
        rng = float(self.data_range) if self.data_range is not None else float(target.max() - target.min())
        mse = torch.mean((preds.double() - target.double()) ** 2)
        return (10.0 * torch.log10(rng ** 2 / mse)).float()


class StructuralSimilarityIndexMeasure(torch.nn.Module):
    def __init__(self, data_range=None, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03, **kwargs):
        super().__init__()
        self.data_range, self.kernel_size, self.sigma, self.k1, self.k2 = data_range, int(kernel_size), float(sigma), k1, k2

    def forward(self, preds, target):
        rng = float(self.data_range) if self.data_range is not None else float(max(preds.max() - preds.min(), target.max() - target.min()))
        c1, c2 = (self.k1 * rng) ** 2, (self.k2 * rng) ** 2
        k, C = self.kernel_size, preds.shape[1]
        d = torch.arange((1 - k) / 2, (1 + k) / 2, 1, dtype=preds.dtype, device=preds.device)
        g = torch.exp(-(d / self.sigma) ** 2 / 2)
        g = (g / g.sum())[:, None]
        kernel = (g @ g.t()).expand(C, 1, k, k).contiguous()
        pad = (k - 1) // 2
        p, t = (F.pad(x, (pad, pad, pad, pad), mode="reflect") for x in (preds, target))
        both = torch.cat([p, t, p * p, t * t, p * t])
        out = F.conv2d(both, kernel, groups=C)
        mu_p, mu_t, pp, tt, pt = out.split(preds.shape[0])
        s_p, s_t, s_pt = pp - mu_p ** 2, tt - mu_t ** 2, pt - mu_p * mu_t
        ssim = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (s_p + s_t + c2))
        ssim = ssim[..., pad:-pad, pad:-pad]
        return ssim.reshape(ssim.shape[0], -1).mean(-1).mean()
