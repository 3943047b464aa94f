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


N_VIEWS, RAW_W, RAW_H = 3, 96, 80


def build_raw_zju(raw_path, processed_path, n_frames=10):
    """A synthetic RAW ZJU-MoCap capture (what reference dataset/test.py reads next to the processed folder):
    ``annots.npy`` {'cams': {'K','R','T' (mm),'D'}}, ``Camera_B<v>/<frame:06d>.jpg``, ``mask/Camera_B<v>/<frame:06d>.png``
    and ``mask_cihp/…`` — plus a processed folder with ``n_frames`` frames for the poses (so the 1/5 novel-pose split is
    not empty)."""
    import os
    from PIL import Image
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.dataset_io import write_synthetic_dataset
    scene = S.make_humanoid(2000, seed=0)
    rng = np.random.default_rng(23)
    poses = S.make_poses(n_frames, seed=7)
    cams = [S.make_camera(azimuth=0.5 * i, img_size=(W, H), focal=60.0, base_size=W) for i in range(n_frames)]
    write_synthetic_dataset(processed_path, scene, poses, cams, (rng.random((n_frames, H, W, 3)) * 255).astype(np.uint8),
                            np.full((n_frames, H, W), 255, np.uint8), Rh=rng.normal(0, 0.3, (n_frames, 3)), Th=rng.normal(0, 0.2, (n_frames, 3)))
    Ks, Rs, Ts, Ds = [], [], [], []
    for v in range(N_VIEWS):
        K, E = S.make_camera(azimuth=2.1 * v + 0.3, img_size=(RAW_W, RAW_H), focal=110.0 + 7 * v, base_size=RAW_W)
        Ks.append(np.asarray(K, np.float64)); Rs.append(np.asarray(E, np.float64)[:3, :3])
        Ts.append(np.asarray(E, np.float64)[:3, 3:4] * 1000.0)                               # millimetres
        Ds.append(DISTORTION[v % len(DISTORTION)].reshape(5, 1))
    np.save(os.path.join(raw_path, "annots.npy"), {"cams": {"K": Ks, "R": Rs, "T": Ts, "D": Ds}}, allow_pickle=True)
    yy, xx = np.mgrid[0:RAW_H, 0:RAW_W]
    smooth = lambda a, b, c: 127 + 120 * np.sin(xx / a + c) * np.cos(yy / b)
    for v in range(N_VIEWS):
        for sub in ("", "mask", "mask_cihp"):
            os.makedirs(os.path.join(raw_path, sub, f"Camera_B{v + 1}"), exist_ok=True)
        for fr in range(n_frames):
            img = np.stack([smooth(7 + v, 9, fr), smooth(11, 5 + v, 2 * fr), smooth(6, 13, v)], -1).clip(0, 255).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(raw_path, f"Camera_B{v + 1}", f"{fr:06d}.jpg"), quality=95)
            m1 = ((xx - 48 - 3 * v) ** 2 / 2 + (yy - 40) ** 2 < (22 + fr) ** 2).astype(np.uint8)
            m2 = ((xx - 40) ** 2 + (yy - 44 - 2 * v) ** 2 < 15 ** 2).astype(np.uint8) * 7        # any non-zero label counts
            Image.fromarray(np.stack([m1] * 3, -1)).save(os.path.join(raw_path, "mask", f"Camera_B{v + 1}", f"{fr:06d}.png"))
            Image.fromarray(np.stack([m2] * 3, -1)).save(os.path.join(raw_path, "mask_cihp", f"Camera_B{v + 1}", f"{fr:06d}.png"))


def write_mdm_motion(path, n_poses, as_torch=True):
    """An MDM-format motion file as reference dataset/newpose.py:152-164 reads it: np.save of a dict with 'thetas_ori'
    [24,3,N] (a torch tensor, as MDM's SMPL export stores it) and 'root_translation' [3,N]."""
    import torch
    rng = np.random.default_rng(31)
    thetas = (rng.normal(0, 0.25, (24, 3, n_poses))).astype(np.float32)
    root = rng.normal(0, 0.3, (3, n_poses)).astype(np.float32)
    np.save(path, {"thetas_ori": torch.from_numpy(thetas) if as_torch else thetas, "root_translation": root}, allow_pickle=True)
    return path
