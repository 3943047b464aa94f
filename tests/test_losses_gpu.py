"""GPU parity: fused photometric losses and LPIPS against the oracle / the reference's own LPIPS (golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as OL

pytestmark = pytest.mark.gpu
t = torch.from_numpy
DEV = "cuda:0"


@pytest.mark.parametrize("layout", ["separate", "rgba_views"])
@pytest.mark.parametrize("with_bg", [True, False])
def test_photometric_l1_forward_backward(layout, with_bg):
    from gomavatar_b200.losses import photometric_l1, unpack
    rng = np.random.default_rng(0)
    B, H, W = 3, 40, 56
    rgba = rng.random((B, H, W, 4)).astype(np.float32)
    bg = rng.random((B, 3)).astype(np.float32) if with_bg else None
    gt, gtm = rng.random((B, H, W, 3)).astype(np.float32), (rng.random((B, H, W)) > 0.5).astype(np.float32)
    g_u = rng.normal(size=(B, H, W, 3)).astype(np.float32)
    # oracle
    o = t(rgba).requires_grad_(True)
    u = OL.unpack(o[..., :3], o[..., 3], t(bg)) if with_bg else o[..., :3]
    l_rgb, l_mask = OL.l1_losses(u, o[..., 3], t(gt), t(gtm))
    ((u * t(g_u)).sum() + 1.0 * l_rgb + 5.0 * l_mask).backward()
    # kernel
    k = t(rgba).to(DEV).requires_grad_(True)
    if layout == "separate":
        rgb_in, mask_in = k[..., :3].contiguous(), k[..., 3].contiguous()
    else:
        rgb_in, mask_in = k[..., :3], k[..., 3]
    ku, kl_rgb, kl_mask = photometric_l1(rgb_in, mask_in, None if bg is None else t(bg).to(DEV), t(gt).to(DEV), t(gtm).to(DEV))
    np.testing.assert_allclose(ku.detach().cpu().numpy(), u.detach().numpy(), atol=1e-7)
    np.testing.assert_allclose(float(kl_rgb), float(l_rgb), rtol=1e-5)
    np.testing.assert_allclose(float(kl_mask), float(l_mask), rtol=1e-5)
    ((ku * t(g_u).to(DEV)).sum() + 1.0 * kl_rgb + 5.0 * kl_mask).backward()
    np.testing.assert_allclose(k.grad.cpu().numpy(), o.grad.numpy(), rtol=1e-5, atol=2e-6)   # d/dmask sums mixed-sign terms
    if with_bg:
        np.testing.assert_allclose(unpack(rgb_in, mask_in, t(bg).to(DEV)).detach().cpu().numpy(), u.detach().numpy(), atol=1e-7)


def test_lpips_matches_reference_golden(golden_dir):
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    g = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    trunk = seeded_random_trunk(0)
    s = float(sum(v.double().abs().sum() for v in trunk.values()))
    if abs(s - float(g["trunk_abs_sum"])) > 1e-6 * s:
        pytest.skip("torchvision's seeded VGG16 init differs from the one the golden was made with")
    ref = g["grad_x0"]
    for precision, tol_val, tol_grad in (("fp32", 1e-4, 1e-3), ("tf32", 2e-2, None)):
        net = LPIPS(trunk, [g[f"lin{k}"] for k in range(5)], conv_precision=precision).to(DEV)
        x0 = t(g["x0"]).to(DEV).requires_grad_(True)
        val = net(2 * x0 - 1, 2 * t(g["x1"]).to(DEV) - 1)
        np.testing.assert_allclose(val.detach().cpu().numpy(), g["value"], rtol=tol_val)
        val.sum().backward()
        rel = np.abs(x0.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        print(f"LPIPS {precision}: value rel err {np.abs(val.detach().cpu().numpy() - g['value']).max() / g['value'].max():.2e}, grad rel err {rel:.2e}")
        if tol_grad is not None:
            assert rel <= tol_grad, (precision, rel)
