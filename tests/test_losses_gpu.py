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


def _grad_close(got, ref, what, max_rel_l2=2e-2, max_abs=5e-2):
    """Gradient check across DIFFERENT convolution implementations (reference CPU fp32 / cuDNN / csrc/conv_first.cu).
    LPIPS is only piecewise smooth: a VGG pre-activation within rounding of zero (observed: 1e-7 vs exactly 0,
    tools/lpips_debug.py) has its ReLU mask decided by the last bit of the convolution, and ONE flipped unit moves the
    input gradient over its whole receptive field (up to 40 x 40 pixels at relu3_3 — a third of these 64 x 64 test
    images) by 1e-3 .. 1e-2 of the maximum; the reference differs from itself by as much between two GPUs.  Across
    implementations the check is therefore a relative L2 error <= 2e-2 and a maximum error <= 5e-2; the strict 1e-3
    of the north star is asserted where it is well defined — on identical activations
    (test_fused_lpips_gradient_matches_torch_autograd_on_same_activations, measured ~1e-5)."""
    d = (got - ref).astype(np.float64)
    rel_l2 = float(np.sqrt((d ** 2).sum() / (ref.astype(np.float64) ** 2).sum()))
    worst = float(np.abs(d).max() / np.abs(ref).max())
    assert rel_l2 <= max_rel_l2 and worst <= max_abs, (what, rel_l2, worst)
    return worst, rel_l2


def test_lpips_matches_reference_golden(golden_dir):
    """The reference's own LPIPS (utils/lpips, run unmodified on the CPU by oracle/make_golden.py) pins value and gradient."""
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    g = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    trunk = seeded_random_trunk(0)
    s = float(sum(v.double().abs().sum() for v in trunk.values()))
    if abs(s - float(g["trunk_abs_sum"])) > 1e-6 * s:
        pytest.skip("torchvision's seeded VGG16 init differs from the one the golden was made with")
    ref = g["grad_x0"]
    for fused in (True, False):
        for precision, tol_val in (("fp32", 1e-4), ("tf32", 2e-2)):
            net = LPIPS(trunk, [g[f"lin{k}"] for k in range(5)], conv_precision=precision, fused=fused).to(DEV)
            x0 = t(g["x0"]).to(DEV).requires_grad_(True)
            val = net(2 * x0 - 1, 2 * t(g["x1"]).to(DEV) - 1)
            np.testing.assert_allclose(val.detach().cpu().numpy(), g["value"], rtol=tol_val)
            val.sum().backward()
            if precision == "fp32":
                worst, rel_l2 = _grad_close(x0.grad.cpu().numpy(), ref, f"fused={fused}")
                print(f"LPIPS fused={fused} fp32: grad max err {worst:.2e} of max, relative L2 error {rel_l2:.2e}")


def test_fused_lpips_gradient_matches_torch_autograd_on_same_activations(golden_dir):
    """Strict 1e-3 (measured ~1e-5): the hand-rolled backward of the fused path against torch autograd over the SAME
    convolution calls on the SAME [pred | gt] batch, so both sides see bit-identical activations and ReLU masks."""
    import torch.nn.functional as F
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk, _TAPS
    g = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    trunk = seeded_random_trunk(0)
    net = LPIPS(trunk, [g[f"lin{k}"] for k in range(5)], conv_precision="fp32", fused=True, conv_epilogue="kernel",
                conv_impl="cudnn").to(DEV)
    net.own_first_conv = False          # same cuDNN convolution as the torch side, so activations are bit-identical
    B = 2
    x1 = t(g["x1"]).to(DEV).permute(0, 2, 3, 1).contiguous()
    k0 = t(g["x0"]).to(DEV).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    wts = torch.tensor([1.0, 0.7], device=DEV)
    kv = net.per_image(k0, x1, from_unit_range=True)
    (kv * wts).sum().backward()
    # torch autograd, same batch composition and the same F.conv2d(no bias) + (bias, ReLU) split
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        r0 = t(g["x0"]).to(DEV).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        x = torch.cat([r0, x1]).permute(0, 3, 1, 2)
        h = ((2 * x - 1) - net.shift) / net.scale
        h = h.contiguous(memory_format=torch.channels_last)
        feats = []
        for i, layer in enumerate(net.features):
            if isinstance(layer, torch.nn.Conv2d):
                h = torch.relu(F.conv2d(h, layer.weight, None, padding=1) + layer.bias[None, :, None, None])
            elif isinstance(layer, torch.nn.MaxPool2d):
                h = layer(h)
            if i in _TAPS:
                feats.append(h)
        tot = 0
        for k, f in enumerate(feats):
            d = (net._unit(f[:B]) - net._unit(f[B:])) ** 2
            tot = tot + (d * getattr(net, f"lin{k}")).sum(dim=1).mean(dim=(1, 2))
        (tot * wts).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    np.testing.assert_allclose(kv.detach().cpu().numpy(), tot.detach().cpu().numpy(), rtol=1e-5)
    ref = r0.grad.cpu().numpy()
    rel = np.abs(k0.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert rel < 1e-3, rel


# ------------------------------------------------------------------------------------- fused LPIPS kernels (csrc/lpips.cu)
def _heads(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    return [g[f"lin{k}"] for k in range(5)]


@pytest.mark.parametrize("C,h,w,pool", [(64, 12, 10, True), (128, 7, 9, True), (256, 6, 6, True), (512, 4, 6, True),
                                        (512, 5, 3, False), (32, 8, 8, True), (64, 37, 50, True), (128, 18, 17, False)])
def test_lpips_tap_kernels_match_torch(C, h, w, pool):
    """One tapped layer: normalise -> squared difference -> 1x1 head -> spatial mean (+ 2x2 max-pool), forward and
    backward incl. ReLU mask and torch's first-maximum tie-breaking, against float64 torch autograd on the CPU."""
    from gomavatar_b200._lib import GomLpipsTapArgs, call, ptr
    import torch.nn.functional as F
    rng = np.random.default_rng(C + h)
    B = 2
    pre = rng.normal(size=(2 * B, h, w, C)).astype(np.float32)
    pre = np.round(pre * 2) / 2                      # many exact ties and exact zeros
    pre[0, :2, :2] = 0.0                             # an all-zero quad (|f| = 0 branch)
    pre[B:, 2:4, 2:4] = pre[:B, 2:4, 2:4]            # pred == gt on a patch
    lin = rng.random(C).astype(np.float32)
    dval = rng.normal(size=B).astype(np.float32)
    ph, pw = h // 2, w // 2
    gp = rng.normal(size=(B, ph, pw, C)).astype(np.float32)
    # float64 reference
    x = t(pre).double().requires_grad_(True)
    a = torch.relu(x)
    unit = lambda f: f / (torch.sqrt((f * f).sum(-1, keepdim=True) + 1e-10) + 1e-10)
    d = ((unit(a[:B]) - unit(a[B:])) ** 2 * t(lin).double()).sum(-1).mean(dim=(1, 2))
    pooled_ref = F.max_pool2d(a.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    loss = (d * t(dval).double()).sum()
    if pool:
        loss = loss + (pooled_ref[:B] * t(gp).double()).sum()
    loss.backward()
    # kernels
    feats = torch.relu(t(pre)).to(DEV).contiguous()
    lin_d, dval_d = t(lin).to(DEV), t(dval).to(DEV)      # named: a temporary would be freed (and reused) before the launch
    sums = torch.zeros(B, device=DEV)
    pooled = torch.empty(2 * B, ph, pw, C, device=DEV) if pool else None
    call("gom_lpips_tap_forward", GomLpipsTapArgs(n_frames=B, height=h, width=w, channels=C, pool=int(pool), feats=ptr(feats),
                                                  lin=ptr(lin_d), layer_sums=ptr(sums), pooled=ptr(pooled)))
    np.testing.assert_allclose(sums.cpu().numpy(), d.detach().numpy(), rtol=2e-5)
    if pool:
        np.testing.assert_array_equal(pooled.cpu().numpy(), pooled_ref.detach().float().numpy())
    g_pre = torch.full((B, h, w, C), float("nan"), device=DEV)
    gpd = t(gp).to(DEV) if pool else None
    call("gom_lpips_tap_backward", GomLpipsTapArgs(n_frames=B, height=h, width=w, channels=C, pool=int(pool), feats=ptr(feats),
                                                   lin=ptr(lin_d), dL_dval=ptr(dval_d), dL_dpooled=ptr(gpd),
                                                   dL_dpre=ptr(g_pre)))
    ref = x.grad[:B].numpy()
    got = g_pre.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_bias_relu_and_relu_backward_kernels():
    from gomavatar_b200._lib import GomBiasReluArgs, GomReluBwdArgs, call, ptr
    rng = np.random.default_rng(5)
    x = rng.normal(size=(3, 5, 7, 64)).astype(np.float32)
    b = rng.normal(size=64).astype(np.float32)
    xd, bd = t(x).to(DEV), t(b).to(DEV)
    call("gom_bias_relu", GomBiasReluArgs(n_pixels=3 * 5 * 7, channels=64, x=ptr(xd), bias=ptr(bd)))
    ref = np.maximum(x + b, 0)
    np.testing.assert_array_equal(xd.cpu().numpy(), ref)
    g = rng.normal(size=x.shape).astype(np.float32)
    gd = t(g).to(DEV)
    call("gom_relu_backward", GomReluBwdArgs(n=g.size, act=ptr(xd), grad=ptr(gd)))
    np.testing.assert_array_equal(gd.cpu().numpy(), g * (ref > 0))


@pytest.mark.parametrize("hw", [(64, 48), (70, 54)])
@pytest.mark.parametrize("impl", ["tcgen05", "cudnn-kernel", "cudnn-cudnn"])
def test_fused_lpips_matches_oracle(hw, impl, golden_dir):
    """Whole loss through the fused path (csrc/conv_first_tc.cu + csrc/conv3x3_tc.cu with 3xTF32 products + csrc/lpips.cu; and,
    as the A/B baseline, cuDNN convolutions with either epilogue) against the CPU oracle, which is itself pinned to the
    reference's LPIPS by tests/test_oracle_golden.py.  Odd sizes exercise ragged convolution tiles and the ragged pooling edge."""
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    H, W = hw
    B = 2
    rng = np.random.default_rng(11)
    x0 = rng.random((B, H, W, 3)).astype(np.float32)
    x1 = np.clip(x0 + rng.normal(0, 0.1, x0.shape), 0, 1).astype(np.float32)
    x1[:, : H // 3] = x0[:, : H // 3]                       # identical background band, as in training frames
    trunk = seeded_random_trunk(0)
    oracle = OL.LPIPSVGG(trunk, _heads(golden_dir))
    o0 = t(x0).requires_grad_(True)
    ov = oracle(2 * o0.permute(0, 3, 1, 2) - 1, 2 * t(x1).permute(0, 3, 1, 2) - 1).reshape(B)
    wts = t(np.array([1.0, -0.5], np.float32))
    (ov * wts).sum().backward()
    conv_impl, _, epilogue = impl.partition("-")
    net = LPIPS(trunk, _heads(golden_dir), conv_precision="fp32", fused=True, conv_epilogue=epilogue or "cudnn", conv_impl=conv_impl).to(DEV)
    k0 = t(x0).to(DEV).requires_grad_(True)
    kv = net.per_image(k0, t(x1).to(DEV), from_unit_range=True)
    np.testing.assert_allclose(kv.detach().cpu().numpy(), ov.detach().numpy(), rtol=1e-4)
    (kv * wts.to(DEV)).sum().backward()
    _grad_close(k0.grad.cpu().numpy(), o0.grad.numpy(), "fused vs CPU oracle")
    # reference-shaped entry point (inputs already in [-1,1], NCHW) agrees with per_image
    v2 = net(2 * t(x0).to(DEV).permute(0, 3, 1, 2) - 1, 2 * t(x1).to(DEV).permute(0, 3, 1, 2) - 1)
    assert v2.shape == (B, 1, 1, 1)
    np.testing.assert_allclose(v2.reshape(B).cpu().numpy(), kv.detach().cpu().numpy(), rtol=1e-5)


@pytest.mark.parametrize("tc", [0, 1])        # 0: FP32-FMA kernels (csrc/conv_first.cu), 1: tcgen05 3xTF32 GEMMs (csrc/conv_first_tc.cu)
@pytest.mark.parametrize("hw", [(5, 3), (70, 54), (64, 130), (9, 200), (300, 301)])
def test_first_convolution_kernels_match_torch_fp32(hw, tc):
    """First LPIPS convolution (3 -> 64, 3x3, pad 1, + bias + ReLU; and its input gradient) against torch's strict-fp32
    convolution.  Sizes cover single / multiple / ragged tiles (8x64 pixel tiles of the FMA kernels, 128-pixel row tiles
    that straddle image rows and images, several tiles per CTA for the tcgen05 kernels)."""
    import torch.nn.functional as F
    from gomavatar_b200._lib import GomConvFirstArgs, call, ptr
    H, W = hw
    N = 3
    g = torch.Generator(device="cpu").manual_seed(H * 7 + W)
    x = torch.randn(N, H, W, 3, generator=g).to(DEV)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(DEV)
    b = torch.randn(64, generator=g).to(DEV)
    go = torch.randn(N, H, W, 64, generator=g).to(DEV)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xr = x.permute(0, 3, 1, 2).double().requires_grad_(True)
        ref = torch.relu(F.conv2d(xr, w.double(), b.double(), padding=1))
        (ref * go.permute(0, 3, 1, 2).double()).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    out = torch.full((N, H, W, 64), float("nan"), device=DEV)
    call("gom_conv_first_forward", GomConvFirstArgs(n_images=N, height=H, width=W, use_tensor_cores=tc, x=ptr(x), weight=ptr(w),
                                                    bias=ptr(b), out=ptr(out)))
    np.testing.assert_allclose(out.cpu().numpy(), ref.permute(0, 2, 3, 1).detach().float().cpu().numpy(), rtol=1e-5, atol=2e-5)
    # the fused path hands over a ReLU-masked gradient; the mask is taken from the float64 reference so that an output
    # within rounding of 0 (a few in 10^6) does not flip it
    gm = (go * (ref.permute(0, 2, 3, 1) > 0)).float().contiguous()
    dx = torch.full((N, H, W, 3), float("nan"), device=DEV)
    scratch = torch.empty(9, N * H * W, 4, device=DEV) if tc else None
    call("gom_conv_first_backward", GomConvFirstArgs(n_images=N, height=H, width=W, use_tensor_cores=tc, weight=ptr(w), dL_dout=ptr(gm),
                                                     dL_dx=ptr(dx), scratch=ptr(scratch)))
    refg = xr.grad.permute(0, 2, 3, 1).float().cpu().numpy()
    np.testing.assert_allclose(dx.cpu().numpy(), refg, rtol=1e-4, atol=1e-5 * np.abs(refg).max())
    if tc:          # fused ReLU backward: unmasked gradient + the convolution's own output
        act = torch.relu(ref).permute(0, 2, 3, 1).float().contiguous()
        dx2 = torch.full((N, H, W, 3), float("nan"), device=DEV)
        call("gom_conv_first_backward", GomConvFirstArgs(n_images=N, height=H, width=W, use_tensor_cores=1, weight=ptr(w), dL_dout=ptr(go),
                                                         dL_dx=ptr(dx2), scratch=ptr(scratch), act=ptr(act)))
        assert torch.equal(dx2, dx)


def test_lpips_frame_groups_on_separate_streams_equal_one_stream(golden_dir):
    """LPIPS(streams=2): the batch is cut into groups of frames that go through the network on their own CUDA streams (forward
    and, through autograd, backward).  Values and the image gradient must equal the single-stream result (per-image math does
    not depend on the grouping), eagerly and when the step is captured in a CUDA graph."""
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    B, H, W = 5, 48, 40
    rng = np.random.default_rng(3)
    x0 = t(rng.random((B, H, W, 3)).astype(np.float32)).to(DEV)
    x1 = t(rng.random((B, H, W, 3)).astype(np.float32)).to(DEV)
    wts = t(rng.normal(size=B).astype(np.float32)).to(DEV)
    res = {}
    for n in (1, 2):
        net = LPIPS(seeded_random_trunk(0), _heads(golden_dir), conv_precision="fp32", streams=n).to(DEV)
        k0 = x0.clone().requires_grad_(True)
        v = net.per_image(k0, x1, from_unit_range=True)
        (v * wts).sum().backward()
        torch.cuda.synchronize()
        res[n] = (v.detach().clone(), k0.grad.clone())
    np.testing.assert_allclose(res[2][0].cpu().numpy(), res[1][0].cpu().numpy(), rtol=1e-6)
    # a group of 2 / 3 images picks other convolution tile shapes than the batch of 5 (other summation order): kink-aware compare
    _grad_close(res[2][1].cpu().numpy(), res[1][1].cpu().numpy(), "2 stream groups vs 1", max_rel_l2=5e-3, max_abs=2e-2)
    # captured: fork / join of the side streams inside the capture
    net = LPIPS(seeded_random_trunk(0), _heads(golden_dir), conv_precision="fp32", streams=2).to(DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        k0 = x0.clone().requires_grad_(True)
        for _ in range(2):
            k0.grad = None
            (net.per_image(k0, x1, from_unit_range=True) * wts).sum().backward()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        k0.grad = None
        with torch.cuda.graph(g, stream=s):
            v = net.per_image(k0, x1, from_unit_range=True)
            (v * wts).sum().backward()
        k0.grad.zero_()
        g.replay()
        torch.cuda.synchronize()
    np.testing.assert_allclose(v.detach().cpu().numpy(), res[1][0].cpu().numpy(), rtol=1e-6)
    # replay vs eager with the same grouping: equal up to the order of the float atomics (K-split reduce-adds, conv1_1 stencil)
    _grad_close(k0.grad.cpu().numpy(), res[2][1].cpu().numpy(), "graph replay vs eager", max_rel_l2=1e-3, max_abs=5e-3)


def test_shade_rgba_matches_torch_formulation():
    """``shade_rgba`` (csrc/photometric.cu::k_shade_fwd/bwd) against the reference's own expression
    ``rgbs = albedos * shadings`` with albedos / masks the channel slices of the rendered image (models/model.py:281-287):
    values exact, gradients to fp32 rounding (the channel sum of dL/dshading is three terms)."""
    from gomavatar_b200.losses import shade_rgba
    rng = np.random.default_rng(5)
    B, H, W = 2, 37, 53
    rgba = rng.random((B, H, W, 4)).astype(np.float32)
    sh = (rng.random((B, H, W, 1)) * 2).astype(np.float32)
    g_rgb, g_m = rng.normal(size=(B, H, W, 3)).astype(np.float32), rng.normal(size=(B, H, W)).astype(np.float32)
    a, s = t(rgba).to(DEV).requires_grad_(True), t(sh).to(DEV).requires_grad_(True)
    ref_rgb, ref_m = a[..., :3] * s, a[..., 3]
    ((ref_rgb * t(g_rgb).to(DEV)).sum() + (ref_m * t(g_m).to(DEV)).sum()).backward()
    a2, s2 = t(rgba).to(DEV).requires_grad_(True), t(sh).to(DEV).requires_grad_(True)
    rgb, m = shade_rgba(a2, s2)
    assert rgb.shape == (B, H, W, 3) and m.shape == (B, H, W)
    assert torch.equal(rgb, ref_rgb.detach()) and torch.equal(m, ref_m.detach())
    ((rgb * t(g_rgb).to(DEV)).sum() + (m * t(g_m).to(DEV)).sum()).backward()
    assert torch.equal(a2.grad, a.grad)
    np.testing.assert_allclose(s2.grad.cpu().numpy(), s.grad.cpu().numpy(), rtol=1e-5, atol=1e-6)
    # only the colour part used (masks without a gradient)
    a3, s3 = t(rgba).to(DEV).requires_grad_(True), t(sh).to(DEV).requires_grad_(True)
    (shade_rgba(a3, s3)[0] * t(g_rgb).to(DEV)).sum().backward()
    assert torch.equal(a3.grad[..., :3], a.grad[..., :3]) and float(a3.grad[..., 3].abs().max()) == 0.0
