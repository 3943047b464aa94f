"""GPU: the per-tile depth sort of csrc/raster_fwd.cu (k_worklist + k_tile_sort) on lists the avatar scenes never produce —
one tile holding 12 000 entries (several counting-sort passes), all depths equal or drawn from two values (every key in one
or two buckets: the long-run and the one-bucket-exceeds-a-pass fallbacks), duplicated depths with different ids (ties must
break by Gaussian index, like upstream's stable radix sort of (tile | depth) keys emitted in index order).
Expected lists are rebuilt from the kernel's own per-Gaussian state (rect, depth — themselves bit-exact against the oracle in
tests/test_raster_gpu.py): for every tile, the ids whose rect covers it, ordered by (depth bits, id)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _expected_lists(rect, depth, radii, gx, gy):
    bits = depth.view(np.uint32).astype(np.uint64)
    out = []
    vis = np.nonzero(radii > 0)[0]
    for ty in range(gy):
        for tx in range(gx):
            m = vis[(rect[vis, 0] <= tx) & (tx < rect[vis, 2]) & (rect[vis, 1] <= ty) & (ty < rect[vis, 3])]
            key = (bits[m] << np.uint64(32)) | m.astype(np.uint64)
            out.append(m[np.argsort(key, kind="stable")].astype(np.uint32))
    return out


def _run(means, cov6, W, H, f=500.0):
    from gomavatar_b200.rasterizer import rasterize_gaussians
    from oracle import camera as Cam
    P = means.shape[0]
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], np.float32)
    st = Cam.raster_settings_from_KE(K, np.eye(4, dtype=np.float32), (W, H))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    aux = {}
    rng = np.random.default_rng(0)
    rasterize_gaussians(t(means)[None], t(cov6)[None], t(rng.random((P, 3)).astype(np.float32)), torch.full((1, P), 0.3, device=DEV),
                        t(st.viewmatrix)[None], t(st.projmatrix)[None], torch.tensor([[st.tanfovx, st.tanfovy]], device=DEV),
                        torch.zeros(1, 3, device=DEV), H, W, aux=aux)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    off = aux["tile_offset"][0].cpu().numpy().view(np.uint32)
    plist = aux["point_list"][0].cpu().numpy().view(np.uint32)
    exp = _expected_lists(aux["rect"][0].cpu().numpy(), aux["depth"][0].cpu().numpy(), aux["radii"][0].cpu().numpy(), gx, gy)
    assert int(aux["status"][0]) == 0
    longest = 0
    for tile in range(T):
        got = plist[off[tile]: off[tile + 1]]
        np.testing.assert_array_equal(got, exp[tile], err_msg=f"tile {tile} ({len(got)} entries)")
        longest = max(longest, len(got))
    wl = aux["worklist"].cpu().numpy().view(np.uint32)
    assert sorted(wl.tolist()) == list(range(T))                          # every tile exactly once ...
    lens = (off[1:T + 1] - off[:T])[wl]
    cls = np.where(lens > 0, np.floor(np.log2(np.maximum(lens, 1)) * 1.0), -1)
    assert np.all(np.diff(cls) <= 0)                                      # ... in non-increasing order of length class
    return longest


def _cov(P, sigma):
    c = np.zeros((P, 6), np.float32)
    c[:, 0] = c[:, 3] = c[:, 5] = sigma * sigma
    return c


@pytest.mark.parametrize("mode", ["random", "same", "two_values", "ties", "one_bucket_over_a_pass"])
def test_one_tile_with_a_very_long_list(mode):
    rng = np.random.default_rng(1)
    P = {"same": 3000, "one_bucket_over_a_pass": 9000}.get(mode, 12000)
    means = np.zeros((P, 3), np.float32)
    means[:, 0] = rng.uniform(-0.02, 0.02, P)                             # all inside the central tiles of a 64 x 64 image
    means[:, 1] = rng.uniform(-0.02, 0.02, P)
    if mode == "random":
        means[:, 2] = rng.uniform(2.0, 6.0, P)
    elif mode in ("same", "one_bucket_over_a_pass"):
        means[:, 2] = 3.0
    elif mode == "two_values":
        means[:, 2] = np.where(rng.random(P) < 0.5, 3.0, 3.0000002)      # adjacent floats: range of one ulp
    else:
        means[:, 2] = rng.choice(rng.uniform(2.0, 6.0, 400).astype(np.float32), P)   # 400 distinct depths, ~30 ids each
    longest = _run(means, _cov(P, 0.004), 64, 64)
    assert longest > (4096 if P > 4096 else 2000)


def test_many_tiles_random_depths_and_sizes():
    rng = np.random.default_rng(2)
    P = 20000
    means = np.stack([rng.uniform(-0.6, 0.6, P), rng.uniform(-0.6, 0.6, P), rng.uniform(1.5, 8.0, P)], 1).astype(np.float32)
    means[::7, 2] = np.round(means[::7, 2] * 4) / 4                       # many exact depth ties
    sig = rng.uniform(0.002, 0.05, P).astype(np.float32)
    c = np.zeros((P, 6), np.float32)
    c[:, 0] = c[:, 3] = c[:, 5] = sig * sig
    _run(means, c, 208, 144, f=300.0)
