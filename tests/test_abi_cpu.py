"""CPU: the C-ABI library builds, loads and exports every symbol include/gom_b200.h declares; reference-facing
argument checking behaves like upstream; the product package never imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from gomavatar_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "gom_b200.h")).read()
    declared = set(re.findall(r"\b(gom_[a-z0-9_A-Z]+)\s*\(", hdr))
    declared = {d for d in declared if not d.endswith("_t")}
    assert len(declared) >= 8
    L = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gom_b200.h but not exported"
    from gomavatar_b200 import _lib
    assert set(_lib.EXPORTS) == declared


def test_binding_struct_sizes_match(built):
    from gomavatar_b200 import _lib
    L = _lib.lib()          # raises on any ABI / struct-size mismatch
    assert L.gom_abi_version() == _lib.ABI_VERSION


def test_invalid_arguments_are_reported_not_thrown(built):
    from gomavatar_b200 import _lib
    L = _lib.lib()
    a = _lib.GomRasterFwdArgs()           # all zero
    assert L.gom_raster_forward(ctypes.byref(a), None) == -1
    assert b"invalid argument" in L.gom_last_error()
    assert L.gom_raster_forward(None, None) == -1


def test_rasterizer_argument_errors_match_upstream():
    from diff_gaussian_rasterization import GaussianRasterizer
    r = GaussianRasterizer(None)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(x, x, torch.ones(4, 1), shs=None, colors_precomp=None, cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.ones(4, 1), colors_precomp=x, scales=None, rotations=None, cov3D_precomp=None)


def test_no_cpu_fallback():
    from gomavatar_b200 import _lib
    from gomavatar_b200.rasterizer import rasterize_gaussians
    z = torch.zeros
    with pytest.raises(_lib.GomError, match="CUDA device"):
        rasterize_gaussians(z(1, 4, 3), z(1, 4, 6), z(4, 3), z(1, 4), z(1, 4, 4), z(1, 4, 4), z(1, 2), z(1, 3), 16, 16)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gomavatar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_rgba_fast_paths_of_the_losses_refuse_cpu_tensors_and_foreign_views():
    """Host logic of gomavatar_b200/losses.py: the fused RGBA paths are taken only for the channel slices of ONE contiguous
    fp32 CUDA [B,H,W,4] tensor (what Model.forward returns); anything else keeps the general path, and CPU tensors are refused
    loudly (there is no CPU path)."""
    import pytest
    import torch
    from gomavatar_b200 import _lib, losses
    rgba = torch.rand(2, 5, 7, 4)
    assert losses._rgba_base(rgba[..., :3], rgba[..., 3]) is None                      # not on a CUDA device
    assert losses._rgba_base(rgba[..., :3].contiguous(), rgba[..., 3]) is None         # not views of one tensor
    other = torch.rand(2, 5, 7, 4)
    assert losses._rgba_base(rgba[..., :3], other[..., 3]) is None
    with pytest.raises(_lib.GomError):
        losses.shade_rgba(rgba, torch.rand(2, 5, 7, 1))
    with pytest.raises(_lib.GomError):
        losses.photometric_l1(rgba[..., :3], rgba[..., 3], None, torch.rand(2, 5, 7, 3), torch.rand(2, 5, 7))
