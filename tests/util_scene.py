"""Shared test helpers: seeded synthetic rasterizer inputs (built with the CPU oracle's geometry)."""
import numpy as np
import torch

from gomavatar_b200 import synthetic as S
from oracle import camera as Cam
from oracle import geometry as G

t = torch.from_numpy
_cache = {}


def raster_inputs(n_faces=2000, img=64, n_frames=1, seed=0, channels=4, focal=537.0, distance=3.5):
    """-> dict(means3D [B,P,3], cov6 [B,P,6], colors [P,C], opacity [B,P], view [B,4,4], proj [B,4,4], tanfov [B,2],
    bg [B,C], H, W) as float32 numpy arrays."""
    key = (n_faces, img, n_frames, seed, channels, focal, distance)
    if key in _cache:
        return _cache[key]
    W, H = (img, img) if np.isscalar(img) else img
    sc = S.make_humanoid(n_faces, seed=seed)
    fr = S.make_frames(sc, n_frames, img_size=(W, H), seed=seed + 3, focal=focal, distance=distance)
    pr = S.make_params(sc, seed=seed + 1)
    means, covs, views, projs, tans = [], [], [], [], []
    for b in range(n_frames):
        _, xyz, cov = G.pose_geometry(t(pr["vertices"]), t(sc.faces), t(sc.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                      t(fr["cnl_gtfms"][b]), t(fr["dst_Rs"][b]), t(fr["dst_Ts"][b]))
        st = Cam.raster_settings_from_KE(fr["K"][b], fr["E"][b], (W, H))
        means.append(xyz.numpy()); covs.append(G.pack_cov6(cov).numpy())
        views.append(st.viewmatrix); projs.append(st.projmatrix); tans.append([st.tanfovx, st.tanfovy])
    app = pr["appearance"].T
    colors = app if channels == 3 else np.concatenate([app, np.ones_like(app[:, :1])], 1)
    rng = np.random.default_rng(seed + 7)
    out = dict(means3D=np.stack(means), cov6=np.stack(covs), colors=np.ascontiguousarray(colors, dtype=np.float32),
               opacity=np.ones((n_frames, n_faces), np.float32), view=np.stack(views), proj=np.stack(projs),
               tanfov=np.asarray(tans, np.float32), bg=rng.uniform(0, 1, (n_frames, channels)).astype(np.float32),
               H=H, W=W, scene=sc, frames=fr, params=pr)
    _cache[key] = out
    return out
