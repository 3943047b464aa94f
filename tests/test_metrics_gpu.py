"""GPU parity: evaluation metrics (reference eval.py:101-108 + utils/image_util.py:21-22) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import losses as OL

pytestmark = pytest.mark.gpu
t = torch.from_numpy
DEV = "cuda:0"


@pytest.mark.parametrize("hw", [(64, 64), (37, 91), (7, 7), (512, 512)])
def test_eval_metrics_match_oracle(hw):
    from gomavatar_b200.metrics import eval_metrics
    H, W = hw
    rng = np.random.default_rng(H * 1000 + W)
    B = 2
    yy, xx = np.mgrid[0:H, 0:W]
    base = 0.5 + 0.4 * np.sin(xx / 7.0)[..., None] * np.cos(yy / 5.0)[..., None] * np.array([1.0, 0.7, -0.5])
    gt = np.clip(base[None] + rng.normal(0, 0.05, (B, H, W, 3)), -0.1, 1.1).astype(np.float32)     # incl. out-of-range values
    pred = (gt + rng.normal(0, 0.03, gt.shape)).astype(np.float32)
    pred[1] = gt[1]                                                   # identical frame: mse 0, psnr inf, ssim 1
    m = eval_metrics(t(pred).to(DEV), t(gt).to(DEV), quantize=True, return_8b=True)
    for b in range(B):
        p8, g8 = OL.to_8b(pred[b]), OL.to_8b(gt[b])
        assert np.array_equal(m["pred_8b"][b].cpu().numpy(), p8)
        p, g = p8 / 255.0, g8 / 255.0
        mse = np.mean((p - g) ** 2)
        assert abs(float(m["mse"][b]) - mse) <= 1e-15 + 1e-12 * mse
        assert abs(float(m["ssim"][b]) - OL.ssim(p, g)) < 1e-11
        if mse > 0:
            assert abs(float(m["psnr"][b]) - OL.psnr(p, g)) < 1e-9
        else:
            assert np.isinf(float(m["psnr"][b]))
    # already-quantised inputs (what Evaluator.evaluate receives) give the same numbers
    q = lambda a: (OL.to_8b(a) / 255.0).astype(np.float32)
    m2 = eval_metrics(t(q(pred)).to(DEV), t(q(gt)).to(DEV), quantize=False)
    assert torch.equal(m2["mse"], m["mse"]) and torch.allclose(m2["ssim"], m["ssim"], rtol=0, atol=1e-13)


def test_evaluator_mirror_accumulates_like_the_reference():
    from gomavatar_b200.metrics import Evaluator
    rng = np.random.default_rng(2)
    ev = Evaluator(lpips_model=None)
    ref = {"psnr": [], "ssim": []}
    for _ in range(3):
        g = OL.to_8b(rng.random((48, 40, 3)).astype(np.float32)) / 255.0
        p = OL.to_8b(np.clip(g + rng.normal(0, 0.05, g.shape), 0, 1).astype(np.float32)) / 255.0
        ev.evaluate(p, g)
        ref["psnr"].append(OL.psnr(p, g)); ref["ssim"].append(OL.ssim(p, g))
    out = ev.summarize()
    assert abs(out["psnr"] - np.mean(ref["psnr"])) < 1e-9 and abs(out["ssim"] - np.mean(ref["ssim"])) < 1e-11
    assert ev.psnr == []
