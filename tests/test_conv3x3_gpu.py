"""GPU parity of csrc/conv3x3_tc.cu (tcgen05 implicit-GEMM 3x3 convolution, forward and dgrad) against float64 torch.

Reference semantics: torchvision VGG16 ``features`` as used by utils/lpips/pretrained_networks.py:96-134 (Conv2d 3x3,
stride 1, padding 1, + ReLU) and its autograd.  Tolerances: 3xTF32 (the strict mode the LPIPS parity tests use) 1e-4
of the output scale; TF32 3e-3 (10-bit mantissa operands — cuDNN's own TF32 kernels measure the same 3e-4 relative L2,
tools/conv_probe.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (N, H, W, C_in, C_out): single / ragged / multiple 16x16 tiles, every accumulator width (32, 64, 128 columns),
# several weight tiles per pixel tile, several channel blocks
CASES = [(2, 20, 24, 64, 64), (1, 8, 16, 32, 32), (3, 17, 33, 64, 128), (1, 40, 48, 128, 256), (2, 9, 50, 96, 160),
         (1, 32, 32, 256, 512), (1, 3, 5, 32, 64), (2, 64, 64, 64, 64)]


def _pack_bits(b):
    n, h, w, c = b.shape
    v = (b.reshape(n, h, w, c // 32, 32).to(torch.int64) << torch.arange(32, device=b.device)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


def _unpack_bits(m, c):
    n, h, w, _ = m.shape
    return ((m.to(torch.int64)[..., None] >> torch.arange(32, device=m.device)) & 1).reshape(n, h, w, c).bool()


def _err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("shape", ["auto", "128,2", "128,1", "64,2", "64,1", "64,1,2", "128,2,4", "64,2,8", "256,1"])
def test_conv3x3_forward_and_dgrad_match_torch(case, precision, shape, monkeypatch):
    """``shape`` pins the kernel's tile shape (output channels per tile, M = 128 sub-tiles per tile[, K-splits: CTAs sharing
    the channel blocks of one tile, partial sums combined by TMA reduce-add + k_conv_finish]); "auto" is the cost model of
    gom_conv3x3.  A shape that does not divide the channel counts falls back (to the automatic choice / to one split)."""
    from gomavatar_b200 import conv as gconv
    if shape != "auto":
        if precision == "fp32" and case[0] * case[1] * case[2] > 3000:
            pytest.skip("3xTF32 on the larger cases is covered by the automatic shape")
        monkeypatch.setenv("GOM_CONV_SHAPE", shape)
    N, H, W, C, K = case
    g = torch.Generator(device="cpu").manual_seed(H * 131 + W * 7 + C)
    x = torch.randn(N, H, W, C, generator=g).relu().to(DEV)
    w = (torch.randn(K, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    b = (torch.randn(K, generator=g) * 0.1).to(DEV)
    go = torch.randn(N, H, W, K, generator=g).to(DEV)
    tol = 3e-3 if precision == "tf32" else 1e-4
    status = torch.zeros(1, dtype=torch.int32, device=DEV)

    ref = torch.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)
    mask = gconv.new_mask(N, H, W, K, DEV)
    out = torch.full((N, H, W, K), float("nan"), device=DEV)
    gconv.conv3x3(x, gconv.pack_weights(w, split=precision == "fp32"), bias=b, relu=True, mask_out=mask, precision=precision,
                  out=out, status=status)
    assert int(status.item()) == 0
    assert torch.isfinite(out).all()
    assert _err(out, ref) < tol
    assert torch.equal(_unpack_bits(mask, K), out > 0)          # the mask is exactly the sign of what was stored

    # dgrad, with the ReLU mask of the layer input fused (mask built from x itself)
    refg = F.conv_transpose2d(go.permute(0, 3, 1, 2).double(), w.double(), None, padding=1).permute(0, 2, 3, 1) * (x > 0)
    gx = torch.full((N, H, W, C), float("nan"), device=DEV)
    gconv.conv3x3(go, gconv.pack_weights(w, transpose=True, split=precision == "fp32"), mask_in=_pack_bits(x > 0),
                  precision=precision, out=gx, status=status)
    assert int(status.item()) == 0
    assert torch.isfinite(gx).all()
    assert _err(gx, refg) < tol
    assert torch.equal(gx == 0, ~(x > 0) | (gx == 0))           # masked entries are exact zeros
    # without a mask: the plain transposed convolution
    gx2 = gconv.conv3x3(go, gconv.pack_weights(w, transpose=True, split=precision == "fp32"), precision=precision, status=status)
    refg2 = F.conv_transpose2d(go.permute(0, 3, 1, 2).double(), w.double(), None, padding=1).permute(0, 2, 3, 1)
    assert _err(gx2, refg2) < tol


def test_conv3x3_is_linear_and_translation_consistent_at_full_size():
    """BASELINE-size property checks (16 x 512 x 512 x 64 does not fit a float64 reference comfortably): the convolution of
    a one-hot image reproduces the packed weights (every tap, every tile position class), and conv(a x) = a conv(x)."""
    from gomavatar_b200 import conv as gconv
    C = K = 64
    g = torch.Generator(device="cpu").manual_seed(3)
    w = (torch.randn(K, C, 3, 3, generator=g) / 24).to(DEV)
    wp = gconv.pack_weights(w)
    wr = wp.cpu()                                                # TF32-rounded weights [9,K,C]
    x = torch.zeros(2, 512, 512, C, device=DEV)
    spots = [(0, 0, 0, 5), (0, 15, 16, 9), (1, 255, 256, 63), (1, 511, 511, 0), (0, 300, 17, 31)]
    for (n, y, xx, c) in spots:
        x[n, y, xx, c] = 1.0
    out = gconv.conv3x3(x, wp)
    expected = torch.zeros(2, 512, 512, K)
    for (n, y, xx, c) in spots:
        for r in range(3):
            for s in range(3):
                yy, xo = y - (r - 1), xx - (s - 1)              # out[p] += x[p + tap - 1] * W[tap]  =>  p = spot - (tap - 1)
                if 0 <= yy < 512 and 0 <= xo < 512:
                    expected[n, yy, xo] += wr[r * 3 + s, :, c]
    assert torch.equal(out.cpu(), expected)                      # one product per output: exact, and zero everywhere else
    xr = torch.randn(2, 512, 512, C, generator=g).to(DEV)
    a = gconv.conv3x3(xr, wp, tma_round=False)
    b2 = gconv.conv3x3(xr * 4.0, wp, tma_round=False)            # a power of two scales TF32 operands exactly
    assert torch.equal(a * 4.0, b2)


def test_conv3x3_rejects_bad_arguments():
    from gomavatar_b200 import conv as gconv
    from gomavatar_b200._lib import GomError
    w = torch.randn(64, 48, 3, 3, device=DEV)
    with pytest.raises(GomError):
        gconv.conv3x3(torch.randn(1, 8, 8, 48, device=DEV), gconv.pack_weights(w))       # 48 input channels: not a multiple of 32
    with pytest.raises(GomError):
        gconv.conv3x3(torch.randn(1, 8, 8, 64), torch.randn(9, 64, 64))                  # CPU tensors
    wp = gconv.pack_weights(torch.randn(64, 64, 3, 3, device=DEV))
    with pytest.raises(GomError):
        gconv.conv3x3(torch.randn(1, 8, 8, 32, device=DEV), wp)                          # channel mismatch
    with pytest.raises(GomError):
        gconv.conv3x3(torch.randn(1, 8, 8, 64, device=DEV), wp, precision="fp32")        # 3xTF32 needs a split pack


@pytest.mark.parametrize("case", [(3, 17, 33, 64, 128), (2, 20, 24, 64, 64), (1, 40, 48, 128, 256), (2, 64, 80, 64, 64), (2, 9, 50, 96, 160)])
@pytest.mark.parametrize("shape", ["128,2", "128,1", "64,2", "64,1", "256,1"])
def test_cta_pair_kernel_equals_single_cta_kernel(case, shape, monkeypatch):
    """k_conv3x3_pair (tcgen05 cta_group::2: two CTAs of a cluster on one M = 256 tile, each with its own halo and half of every
    weight tile) accumulates every output element over the same sequence of MMAs as k_conv3x3, so outputs and ReLU masks must
    be BIT-identical — also where the second CTA's pixel tile lies partly or wholly outside the image (W = 33, 24, 50) and
    with 3xTF32.  GOM_CONV_PAIR=0 selects the single-CTA kernel."""
    from gomavatar_b200 import conv as gconv
    N, H, W, C, K = case
    if K % int(shape.split(",")[0]):
        pytest.skip("tile shape does not divide the channel count")
    monkeypatch.setenv("GOM_CONV_SHAPE", shape)
    g = torch.Generator(device="cpu").manual_seed(H * 31 + W)
    x = torch.randn(N, H, W, C, generator=g).relu().to(DEV)
    w = (torch.randn(K, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    b = (torch.randn(K, generator=g) * 0.1).to(DEV)
    res = {}
    for pair in ("0", "1"):
        monkeypatch.setenv("GOM_CONV_PAIR", pair)
        for precision in ("tf32", "fp32"):
            status = torch.zeros(1, dtype=torch.int32, device=DEV)
            mask = gconv.new_mask(N, H, W, K, DEV)
            out = torch.full((N, H, W, K), float("nan"), device=DEV)
            gconv.conv3x3(x, gconv.pack_weights(w, split=precision == "fp32"), bias=b, relu=True, mask_out=mask, precision=precision,
                          out=out, status=status)
            assert int(status.item()) == 0
            res[pair, precision] = (out, mask)
    for precision in ("tf32", "fp32"):
        assert torch.equal(res["0", precision][0], res["1", precision][0])
        assert torch.equal(res["0", precision][1], res["1", precision][1])
