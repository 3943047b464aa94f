"""CPU: the per-face math shared with the kernels (csrc/gom_face.cuh, compiled for the host as test
infrastructure) against the oracle: forward values, and hand-derived backward vs float64 autograd."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from gomavatar_b200 import synthetic as S
from oracle import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    src = os.path.join(ROOT, "tests", "host_harness", "face_math_host.cpp")
    out = os.path.join(ROOT, "tests", "_build", "libface_math_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out], check=True)
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _case(so3_zero=False):
    sc = S.make_humanoid(2000, seed=2)
    pr = S.make_params(sc, seed=4)
    tri = sc.vertices[sc.faces]                      # [F,3,3]
    so3 = np.ascontiguousarray(pr["so3"].T)
    scale = np.ascontiguousarray(pr["scale"].T)
    if so3_zero:
        so3[:] = 0.0                                 # reference init (model.py:75-85) ...
        scale[::2] = 1.0                             # ... half of the faces isotropic, where d/dso3 == 0 exactly
    else:
        so3[::7] *= 0.05                             # some below the 1e-2 angle clamp
    return tri, so3, scale


@pytest.mark.parametrize("so3_zero", [False, True])
def test_face_forward_and_backward_match_oracle(host, so3_zero):
    tri, so3, scale = _case(so3_zero)
    F = tri.shape[0]
    v = [np.ascontiguousarray(tri[:, k]) for k in range(3)]
    mean, cov6 = np.zeros((F, 3), np.float32), np.zeros((F, 6), np.float32)
    host.face_fwd_host(F, _p(v[0]), _p(v[1]), _p(v[2]), _p(so3), _p(scale), ctypes.c_float(1e-3), _p(mean), _p(cov6))

    def oracle(dtype):
        t = lambda a: torch.tensor(a, dtype=dtype, requires_grad=True)
        tv, tso3, tscale = t(tri), t(so3), t(scale)
        xyz = tv.mean(dim=1)
        Sm = torch.diag_embed(tscale)
        R = G.so3_exp_map(tso3)
        A = G.steiner_frame(tv, 1e-3)
        cov = A @ (R @ Sm @ Sm.transpose(1, 2) @ R.transpose(1, 2)) @ A.transpose(1, 2)
        return tv, tso3, tscale, xyz, G.pack_cov6(cov)

    _, _, _, xyz32, c32 = oracle(torch.float32)
    np.testing.assert_allclose(mean, xyz32.detach().numpy(), atol=1e-6)
    cmax = np.abs(c32.detach().numpy()).max(axis=1, keepdims=True)
    assert (np.abs(cov6 - c32.detach().numpy()) / cmax).max() < 1e-5

    rng = np.random.default_rng(0)
    dmean = rng.normal(size=(F, 3)).astype(np.float32)
    dcov = (rng.normal(size=(F, 6)) * 1e3).astype(np.float32)
    outs = [np.zeros((F, 3), np.float32) for _ in range(5)]
    host.face_bwd_host(F, _p(v[0]), _p(v[1]), _p(v[2]), _p(so3), _p(scale), ctypes.c_float(1e-3), _p(dmean), _p(dcov),
                       *[_p(o) for o in outs])
    tv, tso3, tscale, xyz, c6 = oracle(torch.float64)
    ((xyz * torch.tensor(dmean, dtype=torch.float64)).sum() + (c6 * torch.tensor(dcov, dtype=torch.float64)).sum()).backward()
    ref_dv = tv.grad.numpy()
    for k in range(3):
        ref = ref_dv[:, k]
        assert np.abs(outs[k] - ref).max() <= 2e-4 * np.abs(ref).max(), k
    for name, got, ref in (("so3", outs[3], tso3.grad.numpy()), ("scale", outs[4], tscale.grad.numpy())):
        assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-5, name
