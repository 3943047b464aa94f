"""GPU: the whole chain trains.  A student started from the reference initialisation is fitted to a teacher's renders
with the reference's loss (L1 rgb + 5 L1 mask + LPIPS) through Model.forward, the fused losses, the hand-written
backward kernels and the arena Adam: loss must fall and PSNR / SSIM rise (examples/train_synthetic.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


@pytest.mark.parametrize("with_lpips", [False, True])
def test_student_fits_teacher(with_lpips):
    import train_synthetic as TS
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = TS.make_problem(2000, 64, 4, dev)
    lp = None
    if with_lpips:
        heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
        lp = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)]).to(dev)
    hist = TS.train(model.train(), frames, tgt_rgb, tgt_mask, 60, lpips=lp, decay_steps=60)
    (_, loss0, psnr0, ssim0), (_, loss1, psnr1, ssim1) = hist[0], hist[-1]
    assert loss1 < 0.6 * loss0, (loss0, loss1)
    assert psnr1 > psnr0 + 2.0 and ssim1 > ssim0, (psnr0, psnr1, ssim0, ssim1)
    assert int(model.last_raster_aux["status"].max()) == 0
    for p in model.parameters():
        assert torch.isfinite(p).all()
