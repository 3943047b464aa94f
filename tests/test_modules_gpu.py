"""GPU: the non-rigid MLP's hidden stack on the tcgen05 GEMM kernel (gomavatar_b200.modules._TcMlpStack -> csrc/conv3x3_tc.cu with
kernel_size 1, 3xTF32) against the plain torch formulation of the same module (cuBLAS fp32 — itself pinned to the reference's
own NonRigidModule by tests/test_modules_cpu.py / golden_modules.npz): offsets and every gradient (weights, biases, vertices
through the positional encoding).  Reference: models/modules/non_rigid_module.py:75-147."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("V,B,depth,skips", [(6001, 2, 6, [4]), (4099, 1, 3, [4]), (5000, 3, 6, [2])])
def test_non_rigid_mlp_on_tensor_cores_matches_torch(V, B, depth, skips):
    from gomavatar_b200 import modules as M
    cfg = {"name": "basic", "condition_code_size": 69, "mlp_width": 128, "mlp_depth": depth, "skips": skips, "multires": 6,
           "i_embed": 0, "kick_in_iter": 0, "full_band_iter": 10}
    torch.manual_seed(V)
    net = M.NonRigidModule(cfg).to(DEV)
    with torch.no_grad():                     # the reference's 1e-5 last layer would hide the hidden stack behind rounding
        net.block_mlps[-1].weight.copy_(torch.randn_like(net.block_mlps[-1].weight) * 0.05)
        for m in net.block_mlps:
            if isinstance(m, torch.nn.Linear):
                m.bias.copy_(torch.randn_like(m.bias) * 0.1)
    xyz = (torch.randn(1, 3, V, device=DEV) * 0.5).requires_grad_(True)
    pose = torch.randn(B, 69, device=DEV) * 0.3
    gout = torch.randn(B, 3, V, device=DEV)
    res = {}
    for tc in (False, True):
        M._TC_MLP = tc
        try:
            net.zero_grad(set_to_none=True)
            xyz.grad = None
            out, _, _ = net(xyz, pose, i_iter=1e7)
            (out * gout).sum().backward()
            res[tc] = (out.detach().clone(), xyz.grad.clone(), {n: p.grad.clone() for n, p in net.named_parameters()})
        finally:
            M._TC_MLP = True
    (o0, gx0, gp0), (o1, gx1, gp1) = res[False], res[True]
    assert _rel(o1 - xyz.detach(), o0 - xyz.detach()) < 2e-5                 # the offsets themselves
    # A hidden unit whose pre-activation is within rounding of zero (a handful of the ~1e7 here) takes the other side of the ReLU
    # under a different summation order, which moves the gradient of THAT vertex only: bound the vertices affected and the L2
    # error, and ask the strict tolerance of everything else.
    per_vertex = (gx1 - gx0).abs().amax(dim=(0, 1)) / gx0.abs().max()
    assert float((per_vertex > 1e-4).float().mean()) < 2e-3, float((per_vertex > 1e-4).float().mean())
    assert float((gx1 - gx0).norm() / gx0.norm()) < 5e-3
    # parameter gradients sum over all rows: each flipped unit (see above) shifts them by ~1 / rows of their size.  Both fp32
    # implementations sit equally far (5e-3) from a float64 evaluation of the module for exactly this reason
    # (tools/mlp_precision_check.py: torch fp32 5.4e-3, tcgen05 3xTF32 5.5e-3 of the largest entry), so 1e-2 is the sharpest
    # statement that is true of either.
    for name in gp0:
        assert _rel(gp1[name], gp0[name]) < 1e-2, name
        assert float((gp1[name] - gp0[name]).norm() / gp0[name].norm()) < 1e-2, name


def test_linear_kernel_matches_torch_at_full_size():
    """One layer at the bench size (8 frames x 15 002 vertices = 120 016 rows): y = relu(x W^T + b) and the masked dgrad."""
    from gomavatar_b200 import conv as C
    R = 120016
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(R, 128, generator=g).to(DEV)
    w = (torch.randn(128, 128, generator=g) / 11).to(DEV)
    b = torch.randn(128, generator=g).to(DEV) * 0.1
    mask = C.new_mask(1, R // 16, 16, 128, DEV).view(R, 4)
    y = C.linear(x, C.pack_weights(w, split=True), bias=b, relu=True, mask_out=mask, precision="fp32")
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    assert _rel(y, ref) < 1e-5
    gy = torch.randn(R, 128, generator=g).to(DEV)
    gx = C.linear(gy, C.pack_weights(w, transpose=True, split=True), mask_in=mask, precision="fp32")
    refg = (gy.double() @ w.double()) * (x.double() @ w.double().t() + b.double() > 0)        # masked by THIS layer's output sign
    # the mask belongs to y (the layer's own output), so apply it to a gradient of the same shape: a square layer
    assert _rel(gx, refg) < 1e-5


@pytest.mark.parametrize("R,N", [(4096, 128), (120272, 128), (5008, 192), (48, 32), (16016, 256), (120272, 192)])
def test_linear_wgrad_kernel_matches_float64(R, N):
    """csrc/wgrad_tc.cu: gw = g^T x over R rows with MN-major operands (3xTF32), rows not a multiple of the 32-row k-block, every
    stage count; the bias gradient = column sums formed by the TF32 split pass.  Reference: autograd of nn.Linear
    (models/modules/non_rigid_module.py:75-147)."""
    from gomavatar_b200 import conv as C
    torch.manual_seed(R + N)
    g = torch.randn(R, 128, device=DEV) * torch.rand(1, 128, device=DEV)
    x = torch.randn(R, N, device=DEV).relu_() + 0.01 * torch.randn(R, N, device=DEV)
    gb = torch.zeros(128, device=DEV)
    g_lo = C.tf32_low_part(g, col_sum=gb)
    x_lo = C.tf32_low_part(x)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    gw = C.linear_wgrad(g, g_lo, x, x_lo, status=status)
    assert int(status[0]) == 0
    ref = g.double().T @ x.double()
    scale = g.double().abs().T @ x.double().abs()                 # what a rounding error of either operand is relative to
    err = ((gw.double() - ref).abs() / scale).max()
    assert float(err) < 2e-6, float(err)                          # TF32 alone would be ~1e-3
    # accumulate into an existing buffer
    gw2 = C.linear_wgrad(g, g_lo, x, x_lo, out=gw.clone(), accumulate=True)
    assert float(((gw2.double() - 2 * ref).abs() / scale).max()) < 4e-6
    assert float((gb.double() - g.double().sum(0)).abs().max() / g.double().abs().sum(0).max()) < 1e-5


@pytest.mark.parametrize("R,C,n_out", [(4096, 128, 3), (120272, 128, 3), (5001, 64, 1), (7000, 256, 4), (4100, 132, 2)])
def test_narrow_linear_kernels_match_float64(R, C, n_out):
    """csrc/narrow_linear.cu (the 128 -> 3 output layer of the non-rigid MLP, reference non_rigid_module.py:112-118) against
    torch in float64: forward, input / weight / bias gradients."""
    from gomavatar_b200.modules import _NarrowLinear
    torch.manual_seed(R + C)
    x = torch.randn(R, C, device=DEV)
    w = torch.randn(n_out, C, device=DEV) * 0.1
    b = torch.randn(n_out, device=DEV)
    g = torch.randn(R, n_out, device=DEV)
    xo, wo, bo = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yo = xo @ wo.T + bo
    (yo * g.double()).sum().backward()
    xk, wk, bk = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yk = _NarrowLinear.apply(xk, wk, bk)
    (yk * g).sum().backward()
    assert _rel(yk.detach(), yo.detach()) < 1e-5
    assert _rel(xk.grad, xo.grad) < 1e-5
    assert _rel(wk.grad, wo.grad) < 2e-5
    # a sum of R signed values: the rounding error scales with sum |g|, not with the (cancelling) result
    assert float((bk.grad.double() - bo.grad).abs().max() / g.double().abs().sum(0).max()) < 2e-6
