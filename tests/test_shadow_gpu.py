"""GPU: the tcgen05 shadow MLP (csrc/shadow_mlp.cu through gom_shadow_mlp_forward) against the float64 oracle
(oracle/shadow_mlp.py, pinned by the reference's own module in golden_modules.npz).

Tolerances: the output is a sigmoid in (0,1) that multiplies the rendered albedo (reference model.py:283-287), so the
north-star's 1e-4 relative bound on RGB is asserted as 1e-5 absolute here (measured ~1e-6: 3xTF32 products, fp32
accumulation); gradients within 1e-3 of the largest entry."""
import os

import numpy as np
import pytest
import torch

from gomavatar_b200 import _lib
from gomavatar_b200.modules import ShadowModule
from gomavatar_b200.shadow import FusedShadowModule
from oracle import shadow_mlp as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cfg(depth=3, multires=6):
    return {"multires": multires, "mlp_width": 128, "mlp_depth": depth, "skips": [depth + 1]}


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(1234)            # module construction draws its Xavier weights from the global generator


def _trained_like(m, seed):
    """the reference initialises the last layer at 1e-5 (output == 0.5 everywhere): perturb the hidden layers and give
    the output layer weights that spread the pre-sigmoid value over a few units, so every layer matters"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in list(m.parameters())[:-2]:
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
        m.block_mlps[-1].weight.copy_(torch.randn(m.block_mlps[-1].weight.shape, generator=g) * 0.15)
        m.block_mlps[-1].bias.fill_(0.1)
    return m


def _wb(m):
    lin = [x for x in m.block_mlps if isinstance(x, torch.nn.Linear)]
    return [x.weight.detach().cpu().numpy() for x in lin], [x.bias.detach().cpu().numpy() for x in lin]


def _normals(n, fg_frac, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (1.0 + 2.0 * torch.rand(n, 1, generator=g))      # |n0 + n1 + n2| <= 3
    x[torch.rand(n, generator=g) >= fg_frac] = 0.0
    return x


def test_golden_inputs_of_the_reference_module(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_modules.npz"))
    m = FusedShadowModule(_cfg()).to(DEV)
    m.load_state_dict({k[len("shadow."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("shadow.")})
    with torch.no_grad():
        out = m(torch.from_numpy(g["shadow_in"]).to(DEV))
    assert out.shape == (2, 500, 1)
    np.testing.assert_allclose(out.cpu().numpy(), g["shadow_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,fg_frac,depth,multires", [(1000, 0.5, 3, 6), (77, 1.0, 3, 6), (130, 0.0, 3, 6), (200_001, 0.37, 3, 6),
                                                      (4096, 0.9, 1, 6), (5000, 0.6, 5, 4), (3000, 0.7, 2, 10)])
def test_forward_matches_oracle(n, fg_frac, depth, multires):
    torch.manual_seed(0)
    m = _trained_like(FusedShadowModule(_cfg(depth, multires)), seed=depth * 10 + multires).to(DEV)
    x = _normals(n, fg_frac, seed=n)
    with torch.no_grad():
        out = m(x.to(DEV)[None])[0, :, 0].cpu().numpy()
    W, b = _wb(m)
    ref = O.shadow_forward(x.numpy(), W, b, multires=multires)[:, 0]
    assert ref.std() > 0.02 or fg_frac == 0.0                         # the test is not degenerate
    assert np.abs(out - ref).max() <= 1e-5, np.abs(out - ref).max()
    m.check_status()
    assert int(m._ws["n_fg"].item()) == int((x != 0).any(dim=1).sum())


def test_many_tiles_per_cta_full_frame_batch():
    """2 frames of 512x512 with ~30 % foreground: ~1 200 tiles over 148 CTAs, the weight pipeline wraps many times."""
    m = _trained_like(FusedShadowModule(_cfg()), seed=5).to(DEV)
    x = _normals(2 * 512 * 512, 0.3, seed=11)
    with torch.no_grad():
        out = m(x.to(DEV).reshape(2, 512 * 512, 3)).reshape(-1).cpu().numpy()
    W, b = _wb(m)
    ref = O.shadow_forward(x.numpy(), W, b)[:, 0]
    assert np.abs(out - ref).max() <= 1e-5


@pytest.mark.parametrize("n,fg_frac,depth", [(3000, 0.6, 3), (50_000, 0.2, 3), (2 * 512 * 512, 0.3, 3), (4000, 0.5, 1), (4000, 0.5, 2),
                                             (100, 0.0, 3)])
def test_backward_matches_oracle(n, fg_frac, depth):
    """data gradients (k_shadow_bwd_data), weight/bias gradients (k_shadow_bwd_weights: split-K in tensor memory over up to
    ~8 tiles per CTA at the largest size) and the background row"""
    m = _trained_like(FusedShadowModule(_cfg(depth)), seed=2).to(DEV)
    x = _normals(n, fg_frac, seed=n + 1)
    g_out = torch.randn(n, generator=torch.Generator().manual_seed(4))
    xg = x.to(DEV).requires_grad_(True)
    out = m(xg[None])[0, :, 0]
    (out * g_out.to(DEV)).sum().backward()
    W, b = _wb(m)
    r_out, r_n, r_W, r_b = O.shadow_forward_backward(x.numpy(), W, b, g_out.numpy())
    assert np.abs(out.detach().cpu().numpy() - r_out[:, 0]).max() <= 1e-5

    def close(a, ref, what):
        err = np.abs(a.detach().cpu().numpy().astype(np.float64) - ref).max()
        assert err <= 1e-3 * np.abs(ref).max() + 1e-9, (what, err, np.abs(ref).max())
    # dL/dnormal is per pixel: a hidden unit whose pre-activation is within rounding of 0 has its ReLU decided by the last
    # bit (fp32 here, float64 in the oracle) and moves that ONE pixel's gradient by up to ~1e-2 of the maximum (seen: 1 unit
    # in 6e7 at the largest size, tools/shadow_debug.py).  Strict bound on all but 1e-4 of the pixels, loose bound on the rest.
    e_n = np.abs(xg.grad.detach().cpu().numpy().astype(np.float64) - r_n).max(axis=1)
    n_bad = int((e_n > 1e-3 * np.abs(r_n).max()).sum())
    assert n_bad <= max(2, 1e-4 * n) and e_n.max() <= 5e-2 * np.abs(r_n).max(), (n_bad, e_n.max(), np.abs(r_n).max())
    lin = [t for t in m.block_mlps if isinstance(t, torch.nn.Linear)]
    for l, layer in enumerate(lin):
        close(layer.weight.grad, r_W[l], f"W{l}")
        close(layer.bias.grad, r_b[l], f"b{l}")


def test_backward_is_deterministic():
    m = _trained_like(FusedShadowModule(_cfg()), seed=2).to(DEV)
    x = _normals(300_000, 0.4, seed=12).to(DEV)
    g_out = torch.randn(300_000, device=DEV, generator=torch.Generator(device=DEV).manual_seed(4))
    res = []
    for _ in range(2):
        for p in m.parameters():
            p.grad = None
        xg = x.clone().requires_grad_(True)
        (m(xg[None])[0, :, 0] * g_out).sum().backward()
        res.append([xg.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_training_deeper_than_three_layers_is_refused():
    m = FusedShadowModule(_cfg(5)).to(DEV)
    with pytest.raises(NotImplementedError):
        m(_normals(1000, 0.5, seed=1).to(DEV)[None])                  # parameters require grad -> a backward would follow
    with torch.no_grad():
        assert m(_normals(1000, 0.5, seed=1).to(DEV)[None]).shape == (1, 1000, 1)


def test_capacity_overflow_regrows_when_strict_and_is_flagged_otherwise():
    x = _normals(20_000, 0.8, seed=3).to(DEV)
    m = _trained_like(FusedShadowModule(_cfg(), capacity=1024), seed=2).to(DEV)
    out = m(x.clone().requires_grad_(True)[None])
    assert m._ws["capacity"] >= int((x != 0).any(dim=1).sum())
    m2 = FusedShadowModule(_cfg(), capacity=1024, strict=False).to(DEV)
    m2.load_state_dict(m.state_dict())
    out2 = m2(x.clone().requires_grad_(True)[None])
    assert torch.equal(out, out2)                                     # the forward never depends on the capacity
    with pytest.raises(_lib.GomError):
        m2.check_status()


def test_same_values_as_the_torch_module():
    """the fused module is a drop-in for modules.ShadowModule (the mirror of the reference's module)"""
    m = _trained_like(FusedShadowModule(_cfg()), seed=9).to(DEV)
    t = ShadowModule(_cfg()).to(DEV)
    t.load_state_dict(m.state_dict())
    x = _normals(10_000, 0.5, seed=8).to(DEV)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            assert (m(x[None]) - t(x[None])).abs().max().item() <= 1e-5
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_forward_and_backward_replay_from_a_cuda_graph():
    """strict=False keeps every status on the device: forward + backward captured once, replayed on new inputs"""
    m = _trained_like(FusedShadowModule(_cfg(), strict=False, capacity=4096), seed=6).to(DEV)
    n = 10_000
    x_static = _normals(n, 0.3, seed=21).to(DEV).requires_grad_(True)
    g_static = torch.randn(n, device=DEV)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):                                            # warm-up: workspaces, cuBLAS-free path
            for p in m.parameters():
                p.grad = None
            x_static.grad = None
            (m(x_static[None])[0, :, 0] * g_static).sum().backward()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        for p in m.parameters():
            p.grad = None
        x_static.grad = None
        with torch.cuda.graph(graph, stream=side):
            out_static = m(x_static[None])[0, :, 0]
            (out_static * g_static).sum().backward()
    torch.cuda.synchronize()
    x_new, g_new = _normals(n, 0.35, seed=22).to(DEV), torch.randn(n, device=DEV)
    with torch.no_grad():
        x_static.copy_(x_new)
        g_static.copy_(g_new)
    graph.replay()
    torch.cuda.synchronize()
    got = [out_static.clone(), x_static.grad.clone()] + [p.grad.clone() for p in m.parameters()]
    m.check_status()
    e = FusedShadowModule(_cfg()).to(DEV)
    e.load_state_dict(m.state_dict())
    xe = x_new.clone().requires_grad_(True)
    oe = e(xe[None])[0, :, 0]
    (oe * g_new).sum().backward()
    ref = [oe.detach(), xe.grad] + [p.grad for p in e.parameters()]
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


def test_backward_after_a_second_forward_is_refused():
    """the saved operand images live in the module's workspace: a stale backward must fail loudly, not return wrong gradients"""
    m = _trained_like(FusedShadowModule(_cfg()), seed=3).to(DEV)
    x1 = _normals(2000, 0.5, seed=1).to(DEV).requires_grad_(True)
    x2 = _normals(2000, 0.5, seed=2).to(DEV).requires_grad_(True)
    o1 = m(x1[None]).sum()
    o2 = m(x2[None]).sum()
    o2.backward()
    with pytest.raises(_lib.GomError):
        o1.backward()
