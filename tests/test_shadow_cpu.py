"""CPU: host logic of the fused shadow MLP (gomavatar_b200/shadow.py) and the numerics its kernel relies on.

* the oracle restatement (oracle/shadow_mlp.py) reproduces the reference's own module (golden_modules.npz);
* ``background_row`` — the closed-form one-row backward shadow.py adds for the background pixels — equals float64
  autograd of the oracle at normal = 0 (value, normal gradient and every parameter gradient);
* 3xTF32 (hi*hi + hi*lo + lo*hi with round-to-nearest splits) carries fp32-GEMM accuracy, plain TF32 does not;
* the 128B-swizzle image offset used by k_shadow_prep is a permutation of the 128x32 tile that keeps 16-byte chunks.
"""
import os

import numpy as np
import torch

from gomavatar_b200.modules import ShadowModule
from gomavatar_b200.shadow import FusedShadowModule, background_row
from oracle import shadow_mlp as O


def _gold_module(golden_dir, cls=ShadowModule):
    g = np.load(os.path.join(golden_dir, "golden_modules.npz"))
    m = cls({"multires": int(g["cfg.shadow_module.multires"]), "mlp_width": int(g["cfg.shadow_module.mlp_width"]),
             "mlp_depth": int(g["cfg.shadow_module.mlp_depth"]), "skips": [4]})
    m.load_state_dict({k[len("shadow."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("shadow.")})
    return m, g


def _wb(m):
    lin = [x for x in m.block_mlps if isinstance(x, torch.nn.Linear)]
    return [x.weight.detach().numpy() for x in lin], [x.bias.detach().numpy() for x in lin]


def test_shadow_oracle_matches_reference_module(golden_dir):
    m, g = _gold_module(golden_dir)
    W, b = _wb(m)
    out = O.shadow_forward(g["shadow_in"], W, b, multires=m.multires)
    np.testing.assert_allclose(out, g["shadow_out"], rtol=1e-5, atol=1e-6)


def test_background_row_matches_float64_autograd(golden_dir):
    """the one-row closed form that shadow.py adds for the background pixels (normal == 0)"""
    torch.manual_seed(3)
    m, _ = _gold_module(golden_dir)
    with torch.no_grad():                                   # the golden module is at its 1e-5 init: make the output layer matter
        m.block_mlps[-1].weight.mul_(2e3)
        m.block_mlps[-1].bias.add_(0.1)
    lin = [x for x in m.block_mlps if isinstance(x, torch.nn.Linear)]
    W, b = _wb(m)
    y0, g_x0, dWs, dbs, dw_out, db_out = background_row([x.weight.detach() for x in lin[:-1]], [x.bias.detach() for x in lin[:-1]],
                                                        lin[-1].weight.detach().reshape(-1), lin[-1].bias.detach(), m.multires)
    r_out, r_n, r_W, r_b = O.shadow_forward_backward(np.zeros((1, 3)), W, b, np.ones(1), multires=m.multires)

    def close(a, ref, what):
        ref = np.asarray(ref)
        err = np.abs(np.asarray(a, dtype=np.float64).reshape(ref.shape) - ref).max()
        assert err <= 1e-4 * np.abs(ref).max() + 1e-10, (what, err, np.abs(ref).max())
    close(y0.numpy(), r_out, "y0")
    close(g_x0.numpy(), r_n, "normal")
    assert np.abs(r_n).max() > 0                            # background normals do get a gradient in the reference
    for l in range(len(dWs)):
        close(dWs[l].numpy(), r_W[l], f"W{l}")
        close(dbs[l].numpy(), r_b[l], f"b{l}")
    close(dw_out.numpy(), r_W[-1], "w_out")
    close(db_out.numpy(), r_b[-1], "b_out")


def _tf32(x):
    """round-to-nearest (ties away, like cvt.rna.tf32.f32) to 10 explicit mantissa bits"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def test_three_term_tf32_split_has_fp32_gemm_accuracy():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((256, 128)).astype(np.float32)
    B = (rng.standard_normal((128, 128)) * 0.2).astype(np.float32)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    ah, bh = _tf32(A), _tf32(B)
    al, bl = _tf32(A - ah), _tf32(B - bh)
    d = lambda x, y: x.astype(np.float64) @ y.astype(np.float64)
    three = d(al, bh) + d(ah, bl) + d(ah, bh)
    one = d(ah, bh)
    fp32 = A @ B
    scale = np.abs(exact).max()
    assert np.abs(three - exact).max() / scale < 2e-6
    assert np.abs(three - exact).max() <= 4 * np.abs(fp32 - exact).max() + 1e-7 * scale
    assert np.abs(one - exact).max() / scale > 1e-4                      # plain TF32 would miss the 1e-4 render tolerance


def test_swizzle_image_offset_is_a_chunk_permutation():
    pos = np.empty((128, 32), np.int64)
    for n in range(128):
        for kl in range(32):
            pos[n, kl] = n * 32 + (((kl >> 2) ^ (n & 7)) << 2) + (kl & 3)
    assert sorted(pos.ravel().tolist()) == list(range(4096))
    assert (pos // 32 == np.arange(128)[:, None]).all()                  # a row stays inside its own 128 bytes
    assert (pos[:, 0::4] % 4 == 0).all() and (np.diff(pos.reshape(128, 8, 4), axis=2) == 1).all()


def test_fused_module_keeps_the_reference_state_dict_and_refuses_cpu(golden_dir):
    m, g = _gold_module(golden_dir, FusedShadowModule)
    assert set(m.state_dict()) == {k[len("shadow."):] for k in g.files if k.startswith("shadow.")}
    try:
        m(torch.zeros(1, 4, 3))
    except Exception as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("FusedShadowModule must not have a CPU path")
    try:
        FusedShadowModule({"multires": 6, "mlp_width": 64, "mlp_depth": 3, "skips": [4]})
    except NotImplementedError:
        pass
    else:
        raise AssertionError("unsupported widths must be refused, not silently emulated")
