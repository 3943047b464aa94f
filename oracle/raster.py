"""Oracle (test infrastructure): ctypes front-end of ``raster_oracle.c`` (CPU splat rasterizer, fwd + bwd).

PARITY UNPINNED at this boundary — see ``oracle/__init__.py`` and the header of ``raster_oracle.c``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "raster_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libraster_oracle.so")
_lib = None


def build(force=False):
    """gcc the C restatement (also called from ``__graft_entry__.build()``)."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.gor_preprocess.restype = c_int64
        _lib.gor_preprocess.argtypes = [c_int, c_int, c_int] + [c_void_p] * 5 + [c_float, c_float] + [c_void_p] * 6
        _lib.gor_bin.restype = c_int
        _lib.gor_bin.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        _lib.gor_blend_forward.restype = c_int
        _lib.gor_blend_forward.argtypes = [c_int, c_int, c_int] + [c_void_p] * 9
        _lib.gor_blend_backward.restype = c_int
        _lib.gor_blend_backward.argtypes = [c_int, c_int, c_int, c_int] + [c_void_p] * 13
        _lib.gor_preprocess_backward.restype = c_int
        _lib.gor_preprocess_backward.argtypes = [c_int, c_int, c_int] + [c_void_p] * 5 + [c_float, c_float] + [c_void_p] * 4
        _lib.gor_blend_margins.restype = c_int
        _lib.gor_blend_margins.argtypes = [c_int, c_int, c_int] + [c_void_p] * 4 + [c_float, c_void_p, c_void_p]
        _lib.gor_num_threads.restype = c_int
    return _lib


def num_threads():
    return int(lib().gor_num_threads())


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def forward(means3D, cov6, colors, opacity, view, proj, tanfovx, tanfovy, bg, H, W):
    """One frame. means3D [P,3], cov6 [P,6], colors [P,C], opacity [P], view/proj [4,4] (row-major as the
    reference passes them: E^T, E^T K_ndc^T), bg [C].  Returns a dict with the image and every intermediate
    the parity tests compare bit-exactly (radii, rect, tiles_touched, depth, point_list, ranges)."""
    L = lib()
    means3D, cov6, colors, opacity = _f32(means3D), _f32(cov6), _f32(colors), _f32(opacity).reshape(-1)
    view, proj, bg = _f32(view).reshape(16), _f32(proj).reshape(16), _f32(bg).reshape(-1)
    P, C = colors.shape
    assert bg.shape[0] >= C
    gx, gy = (W + 15) // 16, (H + 15) // 16
    radii = np.zeros(P, np.int32)
    depth = np.zeros(P, np.float32)
    xy = np.zeros((P, 2), np.float32)
    conic_opacity = np.zeros((P, 4), np.float32)
    rect = np.zeros((P, 4), np.int32)
    tiles = np.zeros(P, np.uint32)
    n_dup = L.gor_preprocess(P, H, W, _p(means3D), _p(cov6), _p(opacity), _p(view), _p(proj), tanfovx, tanfovy,
                             _p(radii), _p(depth), _p(xy), _p(conic_opacity), _p(rect), _p(tiles))
    n_dup = int(n_dup)
    keys = np.zeros(max(n_dup, 1), np.uint64)
    plist = np.zeros(max(n_dup, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    rc = L.gor_bin(P, H, W, _p(radii), _p(depth), _p(rect), n_dup, _p(keys), _p(plist), _p(ranges))
    assert rc == 0, rc
    out = np.zeros((C, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    rc = L.gor_blend_forward(H, W, C, _p(plist), _p(ranges), _p(xy), _p(conic_opacity), _p(colors), _p(bg),
                             _p(out), _p(final_T), _p(n_contrib))
    assert rc == 0, rc
    return dict(color=out, final_T=final_T, n_contrib=n_contrib, radii=radii, depth=depth, xy=xy,
                conic_opacity=conic_opacity, rect=rect, tiles_touched=tiles, n_dup=n_dup,
                keys=keys[:n_dup], point_list=plist[:n_dup], ranges=ranges,
                _in=dict(means3D=means3D, cov6=cov6, colors=colors, opacity=opacity, view=view, proj=proj,
                         tanfovx=float(tanfovx), tanfovy=float(tanfovy), bg=bg, H=H, W=W))


def margins(fwd, thr=3e-5):
    """Decision margins of the forward blend (see gor_blend_margins): (margin [H,W] float32 — smallest relative distance of
    any alpha >= 1/255 / T < 1e-4 / power > 0 decision of the pixel to its threshold; fragile [P] bool — Gaussians that
    contribute at a pixel whose margin is below ``thr``).  A CUDA kernel whose exp differs in the last bits may legitimately
    differ from the oracle exactly there, and nowhere else."""
    L = lib()
    i = fwd["_in"]
    P = i["colors"].shape[0]
    H, W = i["H"], i["W"]
    m = np.full((H, W), np.inf, np.float32)
    fr = np.zeros(P, np.uint8)
    plist = np.ascontiguousarray(fwd["point_list"]) if fwd["n_dup"] else np.zeros(1, np.uint32)
    rc = L.gor_blend_margins(P, H, W, _p(plist), _p(fwd["ranges"]), _p(fwd["xy"]), _p(fwd["conic_opacity"]), float(thr), _p(m), _p(fr))
    assert rc == 0, rc
    return m, fr.astype(bool)


def backward(fwd, dL_dcolor):
    """dL_dcolor [C,H,W] -> grads dict (means3D [P,3], cov6 [P,6], colors [P,C], opacity [P], means2D [P,2],
    conic [P,3])."""
    L = lib()
    i = fwd["_in"]
    P, C = i["colors"].shape
    H, W = i["H"], i["W"]
    g = _f32(dL_dcolor).reshape(C, H, W)
    d_mean2D = np.zeros((P, 2), np.float32)
    d_conic = np.zeros((P, 3), np.float32)
    d_op = np.zeros(P, np.float32)
    d_col = np.zeros((P, C), np.float32)
    plist = np.ascontiguousarray(fwd["point_list"]) if fwd["n_dup"] else np.zeros(1, np.uint32)
    rc = L.gor_blend_backward(P, H, W, C, _p(plist), _p(fwd["ranges"]), _p(fwd["xy"]), _p(fwd["conic_opacity"]),
                              _p(i["colors"]), _p(i["bg"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(g),
                              _p(d_mean2D), _p(d_conic), _p(d_op), _p(d_col))
    assert rc == 0, rc
    d_means3D = np.zeros((P, 3), np.float32)
    d_cov6 = np.zeros((P, 6), np.float32)
    rc = L.gor_preprocess_backward(P, H, W, _p(i["means3D"]), _p(i["cov6"]), _p(fwd["radii"]), _p(i["view"]),
                                   _p(i["proj"]), i["tanfovx"], i["tanfovy"], _p(d_mean2D), _p(d_conic),
                                   _p(d_means3D), _p(d_cov6))
    assert rc == 0, rc
    return dict(means3D=d_means3D, cov6=d_cov6, colors=d_col, opacity=d_op, means2D=d_mean2D, conic=d_conic)
