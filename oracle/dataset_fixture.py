"""Test infrastructure: the small seeded dataset that both oracle/make_golden.py::dataset_golden (which reads it with the
reference's own ``dataset/train.py::Dataset``) and tests/test_dataset_cpu.py (which reads it with
``gomavatar_b200.dataset_io.Dataset``) write to disk with ``gomavatar_b200.dataset_io.write_synthetic_dataset``."""
from __future__ import annotations

import numpy as np

W, H, N_FRAMES = 48, 40, 3


def build(path):
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.dataset_io import write_synthetic_dataset
    scene = S.make_humanoid(2000, seed=0)
    rng = np.random.default_rng(11)
    poses = S.make_poses(N_FRAMES, seed=5)                                    # [N,72]
    cams = [S.make_camera(azimuth=0.7 * i, img_size=(W, H), focal=60.0 + 3 * i, base_size=W) for i in range(N_FRAMES)]
    yy, xx = np.mgrid[0:H, 0:W]
    images = (rng.random((N_FRAMES, H, W, 3)) * 255).astype(np.uint8)
    masks = np.stack([(((xx - W / 2 - 2 * i) ** 2 + (yy - H / 2) ** 2) < (10 + i) ** 2).astype(np.uint8) * 255 for i in range(N_FRAMES)])
    masks[0, 5:8, 5:9] = 128                                                  # a soft edge value survives the /255 path
    Rh = rng.normal(0, 0.3, (N_FRAMES, 3))
    Th = rng.normal(0, 0.2, (N_FRAMES, 3))
    write_synthetic_dataset(path, scene, poses, cams, images, masks, Rh=Rh, Th=Th)
    return scene, poses, cams, images, masks


DISTORTION = np.array([[0.0, 0.0, 0.0, 0.0, 0.0], [-0.12, 0.05, 0.002, -0.001, 0.01], [0.2, -0.08, -0.003, 0.002, 0.0]])


def add_distortion(path):
    """Rewrites cameras.pkl with non-zero lens distortion for frames 1 and 2 (ZJU-MoCap's processed folders carry the raw
    coefficients; the reader undistorts with them: reference dataset/train.py:149-153)."""
    import os
    import pickle
    with open(os.path.join(path, "cameras.pkl"), "rb") as f:
        cams = pickle.load(f)
    for i, name in enumerate(sorted(cams)):
        cams[name]["distortions"] = DISTORTION[i % len(DISTORTION)].copy()
    with open(os.path.join(path, "cameras.pkl"), "wb") as f:
        pickle.dump(cams, f)
