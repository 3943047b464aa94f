"""CPU: self-consistency of the rasterizer oracle (PARITY UNPINNED vs upstream — see oracle/__init__.py).

* the C oracle's hand-written backward == autograd of an independent float64 torch restatement
* structural invariants of App. A (mask = 1 - final_T, sorted lists, sum(tiles_touched) = N_dup, ...)
"""
import numpy as np
import pytest
import torch

from gomavatar_b200 import synthetic as S
from oracle import camera as Cam
from oracle import geometry as G
from oracle import raster as R
from oracle import raster_torch as RT

t = torch.from_numpy


def _scene(n_faces=2000, img=64, b=1):
    sc = S.make_humanoid(n_faces)
    fr = S.make_frames(sc, 2, img_size=img)
    pr = S.make_params(sc)
    _, xyz, cov = G.pose_geometry(t(pr["vertices"]), t(sc.faces), t(sc.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                  t(fr["cnl_gtfms"][b]), t(fr["dst_Rs"][b]), t(fr["dst_Ts"][b]))
    st = Cam.raster_settings_from_KE(fr["K"][b], fr["E"][b], (img, img))
    colors = np.concatenate([pr["appearance"].T, np.ones((sc.n_faces, 1), np.float32)], 1)
    return xyz.numpy(), G.pack_cov6(cov).numpy(), colors, np.ones(sc.n_faces, np.float32), st


@pytest.fixture(scope="module")
def small():
    xyz, cov6, colors, op, st = _scene()
    bg = np.array([0.1, 0.2, 0.3, 0.0], np.float32)
    f = R.forward(xyz, cov6, colors, op, st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, bg, 64, 64)
    return xyz, cov6, colors, op, st, bg, f


def test_invariants(small):
    xyz, cov6, colors, op, st, bg, f = small
    assert f["n_dup"] == int(f["tiles_touched"].sum()) == len(f["point_list"])
    assert np.abs(f["color"][3] - (1 - f["final_T"])).max() < 1e-6          # ones channel over bg 0 == mask
    keys = f["keys"]
    assert np.all(keys[1:] >= keys[:-1])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for tile in np.unique(tiles):
        s, e = f["ranges"][tile]
        assert np.all(tiles[s:e] == tile) and (s == 0 or tiles[s - 1] != tile)
        d = f["depth"][f["point_list"][s:e]]
        assert np.all(np.diff(d) >= 0)
    rect = f["rect"]
    assert np.array_equal((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1]), f["tiles_touched"].astype(np.int32))
    assert f["radii"].max() > 0 and (f["radii"] >= 0).all()


def test_near_cull_and_offscreen():
    xyz, cov6, colors, op, st = _scene()
    xyz = xyz.copy()
    view = st.viewmatrix.reshape(4, 4)
    z = (np.concatenate([xyz, np.ones((len(xyz), 1), np.float32)], 1) @ view)[:, 2]
    # push half of the Gaussians behind the near plane / far off screen
    xyz[::2] += (0.15 - z[::2])[:, None] * view[:3, 2][None] / np.dot(view[:3, 2], view[:3, 2])
    f = R.forward(xyz, cov6, colors, op, st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, np.zeros(4, np.float32), 64, 64)
    assert np.all(f["radii"][::2] == 0) and np.all(f["tiles_touched"][::2] == 0)
    assert f["radii"][1::2].max() > 0


def test_backward_is_derivative_of_forward(small):
    xyz, cov6, colors, op, st, bg, f = small
    rng = np.random.default_rng(0)
    dL = rng.normal(size=(4, 64, 64)).astype(np.float32)
    g = R.backward(f, dL)
    m = t(xyz).double().requires_grad_(True)
    c6 = t(cov6).double().requires_grad_(True)
    col = t(colors).double().requires_grad_(True)
    o = t(op).double().requires_grad_(True)
    img, _ = RT.render(m, c6, col, o, t(st.viewmatrix), t(st.projmatrix), st.tanfovx, st.tanfovy, t(bg), 64, 64)
    assert np.abs(img.detach().numpy() - f["color"]).max() < 1e-5
    (img * t(dL).double()).sum().backward()
    for name, ref in (("means3D", m.grad), ("cov6", c6.grad), ("colors", col.grad), ("opacity", o.grad)):
        ref = ref.numpy()
        err = np.abs(ref - g[name]).max() / np.abs(ref).max()
        assert err < 1e-4, (name, err)


def test_ragged_image_and_empty_input():
    xyz, cov6, colors, op, st = _scene(img=64)
    # 50x70 image: not a multiple of the 16-px tile
    f = R.forward(xyz, cov6, colors, op, st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, np.zeros(4, np.float32), 50, 70)
    assert f["color"].shape == (4, 50, 70) and f["ranges"].shape[0] == 4 * 5
    # nothing visible -> background only
    f0 = R.forward(xyz + 100.0, cov6, colors, op, st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy,
                   np.array([0.5, 0.25, 0.125, 0.0], np.float32), 64, 64)
    assert f0["n_dup"] == 0 and np.all(f0["color"][0] == 0.5) and np.all(f0["final_T"] == 1.0)
    g0 = R.backward(f0, np.ones((4, 64, 64), np.float32))
    assert all(np.all(v == 0) for v in g0.values())
