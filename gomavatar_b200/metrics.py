"""Evaluation metrics of the reference on libgom_b200.so: PSNR and SSIM exactly as ``eval.py::Evaluator`` defines them
(eval.py:101-108, skimage 0.18 defaults, on 8-bit-quantised images: utils/image_util.py:21-22, eval.py:355-361) plus
LPIPS x 1000 (eval.py:110-116).  One kernel launch for any number of frames, results stay on the device."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomEvalMetricsArgs, call, ptr


def eval_metrics(pred, gt, quantize=True, return_8b=False):
    """pred, gt: [B,H,W,3] float images.  quantize=True applies ``to_8b_image`` first (raw network outputs, the path of
    eval.py:355-361); quantize=False expects images that already are k/255.  Returns a dict of float64 tensors [B]:
    mse, psnr, ssim (and 'pred_8b' uint8 [B,H,W,3] on request)."""
    if pred.device.type != "cuda":
        raise _lib.GomError("eval_metrics: inputs must live on a CUDA device (no CPU path exists)")
    if pred.dim() == 3:
        pred, gt = pred[None], gt[None]
    B, H, W, C = pred.shape
    if C != 3 or gt.shape != pred.shape:
        raise ValueError("eval_metrics expects matching [B,H,W,3] images")
    p, g = pred.detach().contiguous().float(), gt.detach().contiguous().float()
    dev = p.device
    ssim_sum = torch.empty(B, dtype=torch.float64, device=dev)
    sq = torch.empty(B, dtype=torch.int64, device=dev)
    p8 = torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev) if return_8b else None
    call("gom_eval_metrics", GomEvalMetricsArgs(n_frames=B, height=H, width=W, quantize=int(bool(quantize)), pred=ptr(p),
                                                gt=ptr(g), ssim_sum=ptr(ssim_sum), sq_err_sum=ptr(sq), pred_8b=ptr(p8)))
    mse = sq.double() / (65025.0 * 3 * H * W)
    out = {"mse": mse, "psnr": -10.0 * torch.log(mse) / torch.log(torch.tensor(10.0, dtype=torch.float64, device=dev)),
           "ssim": ssim_sum / (3.0 * (H - 6) * (W - 6))}
    if return_8b:
        out["pred_8b"] = p8
    return out


class Evaluator:
    """Mirror of reference ``eval.py::Evaluator`` (:86-143): ``evaluate(rgb_pred, rgb_gt)`` takes [H,W,3] images in [0,1]
    that have already been 8-bit quantised (numpy or torch), appends mse / psnr / ssim / lpips, ``summarize`` returns the
    means.  ``lpips_model`` is a ``gomavatar_b200.lpips.LPIPS`` (or None to skip LPIPS)."""

    def __init__(self, lpips_model=None, device="cuda:0"):
        self.lpips_model, self.device = lpips_model, torch.device(device)
        self.mse, self.psnr, self.ssim, self.lpips = [], [], [], []

    def evaluate(self, rgb_pred, rgb_gt):
        p = torch.as_tensor(rgb_pred).to(self.device).float()[None]
        g = torch.as_tensor(rgb_gt).to(self.device).float()[None]
        m = eval_metrics(p, g, quantize=False)
        self.mse.append(float(m["mse"][0])); self.psnr.append(float(m["psnr"][0])); self.ssim.append(float(m["ssim"][0]))
        if self.lpips_model is not None:
            with torch.no_grad():
                v = self.lpips_model(p.permute(0, 3, 1, 2) * 2. - 1., g.permute(0, 3, 1, 2) * 2. - 1.)
            self.lpips.append(float(v.mean()) * 1000)

    def evaluate_batch(self, rgb_pred, rgb_gt, return_8b=False):
        """The batched form of the loop at eval.py:346-366: raw network outputs / targets ``[B,H,W,3]`` on the device;
        ``to_8b_image`` of both (eval.py:355,361), mse / psnr / ssim of every frame in ONE launch, LPIPS on the quantised
        images.  Returns the quantised predictions (uint8, what eval.py writes to PNG) on request."""
        p, g = rgb_pred.to(self.device).float(), rgb_gt.to(self.device).float()
        m = eval_metrics(p, g, quantize=True, return_8b=True)
        self.mse += m["mse"].tolist(); self.psnr += m["psnr"].tolist(); self.ssim += m["ssim"].tolist()
        if self.lpips_model is not None:
            g8 = (255.0 * g.clamp(0, 1)).to(torch.uint8)
            with torch.no_grad():
                v = self.lpips_model((m["pred_8b"].float() / 255.0).permute(0, 3, 1, 2) * 2. - 1., (g8.float() / 255.0).permute(0, 3, 1, 2) * 2. - 1.)
            self.lpips += (v.reshape(-1) * 1000).tolist()
        return m["pred_8b"] if return_8b else None

    def summarize(self, path=None):
        """Means of the collected metrics; with ``path`` also the reference's result file (eval.py:131-143: ``np.save`` of
        ``{'mse': [...], 'psnr': [...], 'ssim': [...], 'lpips': [...]}``, read back with ``allow_pickle=True``)."""
        mean = lambda v: float(sum(v) / len(v)) if v else float("nan")
        if path is not None:
            import os
            import numpy as np
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            np.save(path, {"mse": list(self.mse), "psnr": list(self.psnr), "ssim": list(self.ssim), "lpips": list(self.lpips)})
        out = {"mse": mean(self.mse), "psnr": mean(self.psnr), "ssim": mean(self.ssim), "lpips": mean(self.lpips)}
        self.mse, self.psnr, self.ssim, self.lpips = [], [], [], []
        return out
