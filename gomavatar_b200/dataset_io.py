"""On-disk compatibility with the reference's processed-dataset format.

* ``write_dataset``   writes what reference ``scripts/prepare_zju-mocap/prepare_dataset.py:143-197`` writes:
                      ``images/<name>.png``, ``masks/<name>.png`` (3-channel), ``cameras.pkl`` {name: intrinsics, extrinsics,
                      distortions}, ``mesh_infos.pkl`` {name: Rh, Th, poses[72], joints[24,3], tpose_joints[24,3]},
                      ``canonical_joints.pkl`` {vertex, joints, weights, edges, faces}, ``avg_betas.npy``.
* ``Dataset``         reads it back with the constructor, ``__getitem__`` keys and ``get_canonical_info()`` of reference
                      ``dataset/train.py::Dataset`` (:18-310), so its items feed ``Model.forward`` / ``compute_loss`` exactly like
                      the reference's (train.py:300-330) and reference-prepared ZJU / Snapshot folders load unchanged.
* ``write_synthetic_dataset``  the synthetic scene of ``gomavatar_b200.synthetic`` in that format (teacher renders as images).

Pinned by the reference's own reader run on folders this module wrote: tests/golden/golden_dataset.npz (every item field,
OpenCV stubbed out) and tests/golden/golden_dataset_cv2.npz (the reference reader with the real OpenCV: lens undistortion —
ZJU-MoCap's processed folders carry the raw coefficients —, LANCZOS4 / LINEAR resampling to a ``target_size``, the
``resize_img_scale`` branch).  Where OpenCV is installed this reader makes the reference's own ``cv2.undistort`` /
``cv2.resize`` calls and the pixels are identical; where it is not, images are resampled with Pillow (close, not identical:
unpinned) and non-zero distortion is refused.  Host-side I/O, not the hot path.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from .synthetic import SMPL_PARENTS, body_pose_to_body_RTs, canonical_global_tfms, rvec_to_rmtx


# ------------------------------------------------------------------------------------------------------------ helpers
def _rodrigues(rvec):
    """cv2.Rodrigues(rvec)[0] (exact axis-angle, no regularisation) in float64."""
    v = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(v))
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * np.outer(k, k)


def apply_global_tfm_to_camera(E, Rh, Th, return_global_tfms=False):
    """reference utils/camera_util.py:111-131."""
    g = np.eye(4)
    rot = _rodrigues(Rh).T
    g[:3, :3] = rot
    g[:3, 3] = -rot.dot(np.asarray(Th, dtype=np.float64))
    out = np.asarray(E).dot(np.linalg.inv(g))
    return (out, g) if return_global_tfms else out


def get_joints_from_pose(poses, tpose_joints):
    """reference utils/body_util.py:553-582 (fp32 forward kinematics of the 24 joints)."""
    N = tpose_joints.shape[0]
    poses = np.asarray(poses).reshape(N, -1)
    T = np.eye(4, dtype=np.float32)[None].repeat(N, axis=0)
    for i in range(N):
        T[i, :3, :3] = rvec_to_rmtx(poses[i]).astype(np.float32)
        T[i, :3, 3] = tpose_joints[i] if i == 0 else tpose_joints[i] - tpose_joints[SMPL_PARENTS[i]]
    joints = np.zeros((N, 4), dtype=np.float32)
    joints[0] = T[0] @ np.array([0, 0, 0, 1], dtype=np.float32)
    for i in range(1, N):
        T[i] = T[SMPL_PARENTS[i]] @ T[i]
        joints[i] = T[i] @ np.array([0, 0, 0, 1], dtype=np.float32)
    return joints[:, :3] / joints[:, 3:]


def unique_edges(faces):
    """trimesh.Trimesh(...).edges of the reference writer is every directed face edge; readers only use it as an index
    list, so the [3F,2] directed edges in face order are written (same content as trimesh's ``edges``)."""
    f = np.asarray(faces)
    return np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=1).reshape(-1, 2)


def _save_png(array_u8, path):
    from PIL import Image
    Image.fromarray(array_u8).save(path)


def _load_rgb(path):
    from PIL import Image
    return np.array(Image.open(path).convert("RGB"))


# ------------------------------------------------------------------------------------------------------------- writer
def write_dataset(path, canonical, frames):
    """canonical: dict(vertex [V,3], joints [24,3], weights [V,24], faces [F,3]); frames: iterable of dicts with keys
    name, image (uint8 [H,W,3]), mask (uint8 [H,W] or [H,W,3], 255 = subject), K [3,3], E [4,4], optional D [5], Rh [3],
    Th [3], poses [72], tpose_joints [24,3], optional joints [24,3] (posed joints; forward kinematics if absent)."""
    os.makedirs(os.path.join(path, "images"), exist_ok=True)
    os.makedirs(os.path.join(path, "masks"), exist_ok=True)
    cameras, mesh_infos = {}, {}
    for fr in frames:
        name = fr["name"]
        mask = np.asarray(fr["mask"], dtype=np.uint8)
        if mask.ndim == 2:
            mask = np.stack([mask] * 3, axis=-1)
        _save_png(np.asarray(fr["image"], dtype=np.uint8), os.path.join(path, "images", f"{name}.png"))
        _save_png(mask, os.path.join(path, "masks", f"{name}.png"))
        cameras[name] = {"intrinsics": np.asarray(fr["K"], dtype=np.float64), "extrinsics": np.asarray(fr["E"], dtype=np.float64),
                         "distortions": np.asarray(fr.get("D", np.zeros(5)), dtype=np.float64)}
        poses = np.asarray(fr["poses"], dtype=np.float64).reshape(-1)
        tpose = np.asarray(fr["tpose_joints"], dtype=np.float64)
        joints = fr.get("joints")
        if joints is None:
            joints = get_joints_from_pose(poses.astype(np.float32), tpose.astype(np.float32))
        mesh_infos[name] = {"Rh": np.asarray(fr.get("Rh", np.zeros(3)), dtype=np.float64),
                            "Th": np.asarray(fr.get("Th", np.zeros(3)), dtype=np.float64), "poses": poses,
                            "joints": np.asarray(joints, dtype=np.float64), "tpose_joints": tpose}
    with open(os.path.join(path, "cameras.pkl"), "wb") as f:
        pickle.dump(cameras, f)
    with open(os.path.join(path, "mesh_infos.pkl"), "wb") as f:
        pickle.dump(mesh_infos, f)
    np.save(os.path.join(path, "avg_betas.npy"), np.zeros(10))
    faces = np.asarray(canonical["faces"])
    with open(os.path.join(path, "canonical_joints.pkl"), "wb") as f:
        pickle.dump({"vertex": np.asarray(canonical["vertex"], dtype=np.float64),
                     "joints": np.asarray(canonical["joints"], dtype=np.float64),
                     "weights": np.asarray(canonical["weights"], dtype=np.float64), "edges": unique_edges(faces), "faces": faces}, f)


def write_synthetic_dataset(path, scene, poses72, cameras, images, masks, Rh=None, Th=None):
    """``scene``: synthetic.Scene; poses72 [N,72]; cameras: list of (K [3,3], E [4,4]) as the MODEL sees them (i.e. after the
    global transform); images uint8 [N,H,W,3], masks uint8 [N,H,W].  With Rh/Th [N,3] the stored extrinsics are E G so
    that the reader's apply_global_tfm_to_camera recovers E."""
    frames = []
    for i, (K, E) in enumerate(cameras):
        rh = np.zeros(3) if Rh is None else np.asarray(Rh[i], dtype=np.float64)
        th = np.zeros(3) if Th is None else np.asarray(Th[i], dtype=np.float64)
        _, g = apply_global_tfm_to_camera(np.eye(4), rh, th, return_global_tfms=True)
        frames.append({"name": f"frame_{i:06d}", "image": images[i], "mask": masks[i], "K": K, "E": np.asarray(E, dtype=np.float64) @ g,
                       "Rh": rh, "Th": th, "poses": poses72[i], "tpose_joints": scene.joints})
    canonical = {"vertex": scene.vertices, "joints": scene.joints, "weights": scene.lbs_weights[:24].T, "faces": scene.faces}
    write_dataset(path, canonical, frames)


# ------------------------------------------------------------------------------------------------------------- reader
class Dataset:
    """Mirror of reference ``dataset/train.py::Dataset`` (same constructor arguments, item keys and dtypes).  Works as a
    ``torch.utils.data.Dataset`` (``__len__`` / ``__getitem__``); it does not import torch."""

    def __init__(self, dataset_path, keyfilter=None, maxframes=-1, bgcolor=None, skip=1, target_size=None,
                 crop_size=(-1, -1), prefetch=False, split_for_pose=False):
        self.cfg = {"bbox_offset": 0.3, "resize_img_scale": [0.5, 0.5]}
        self.dataset_path = dataset_path
        self.image_dir = os.path.join(dataset_path, "images")
        with open(os.path.join(dataset_path, "canonical_joints.pkl"), "rb") as f:
            cj = pickle.load(f)
        self.canonical_joints = cj["joints"].astype("float32")
        self.canonical_bbox = self.skeleton_to_bbox(self.canonical_joints)
        self.canonical_vertex = cj["vertex"].astype("float32")
        self.canonical_lbs_weights = cj["weights"].astype("float32")
        self.edges = cj["edges"].astype(int) if "edges" in cj else None
        self.faces = cj.get("faces")
        with open(os.path.join(dataset_path, "cameras.pkl"), "rb") as f:
            self.cameras = pickle.load(f)
        with open(os.path.join(dataset_path, "mesh_infos.pkl"), "rb") as f:
            self.mesh_infos = pickle.load(f)
        for name in self.mesh_infos:
            self.mesh_infos[name]["bbox"] = self.skeleton_to_bbox(self.mesh_infos[name]["joints"])
        names = sorted(os.path.splitext(n)[0] for n in os.listdir(self.image_dir) if n.endswith(".png"))
        self.framelist = names[::skip]
        if maxframes > 0:
            self.framelist = self.framelist[:maxframes]
        if split_for_pose:
            self.framelist = self.framelist[:-(len(self.framelist) // 5)]
        self.keyfilter, self.bgcolor = keyfilter, bgcolor
        if target_size is not None:
            self.cfg["target_size"] = target_size
        self.cfg["crop_size"] = list(crop_size)
        self.prefetch = prefetch
        if prefetch:
            self.preload = {n: list(self.load_image(n, np.zeros(3, dtype="float32"))) for n in self.framelist}

    def skeleton_to_bbox(self, skeleton):
        return {"min_xyz": np.min(skeleton, axis=0) - self.cfg["bbox_offset"], "max_xyz": np.max(skeleton, axis=0) + self.cfg["bbox_offset"]}

    @staticmethod
    def _cv2():
        """OpenCV if it is installed (then pixels are produced by the very calls the reference makes), else None."""
        try:
            import cv2
            return cv2
        except Exception:
            return None

    def _resize(self, arr, size, lanczos, scale=None):
        """reference dataset/train.py:157-173: ``cv2.resize`` with INTER_LANCZOS4 (image) / INTER_LINEAR (mask), to
        ``size`` = (w, h) or by ``scale`` = (fx, fy).  Without OpenCV: Pillow's LANCZOS / BILINEAR (close, not identical)."""
        if scale is None and arr.shape[1] == int(size[0]) and arr.shape[0] == int(size[1]):
            return arr                           # cv2.resize to the same size is an exact copy: skip the library entirely
        cv2 = self._cv2()
        if cv2 is not None:
            interp = cv2.INTER_LANCZOS4 if lanczos else cv2.INTER_LINEAR
            if scale is not None:
                return cv2.resize(arr, None, fx=scale[0], fy=scale[1], interpolation=interp)
            return cv2.resize(arr, [int(size[0]), int(size[1])], interpolation=interp)
        from PIL import Image
        if scale is not None:
            size = (int(round(arr.shape[1] * scale[0])), int(round(arr.shape[0] * scale[1])))
        w, h = size
        if arr.shape[1] == w and arr.shape[0] == h:
            return arr
        mode = Image.LANCZOS if lanczos else Image.BILINEAR
        chans = [np.asarray(Image.fromarray(arr[..., c].astype(np.float32), mode="F").resize((w, h), mode)) for c in range(arr.shape[2])]
        return np.stack(chans, axis=-1).astype(arr.dtype)

    def load_image(self, frame_name, bg_color):
        orig = _load_rgb(os.path.join(self.image_dir, f"{frame_name}.png"))
        orig_H, orig_W, _ = orig.shape
        alpha = _load_rgb(os.path.join(self.dataset_path, "masks", f"{frame_name}.png"))
        cam = self.cameras.get(frame_name, {})
        # reference dataset/train.py:149-153; cv2.undistort with all-zero coefficients returns its input bit for bit
        # (tests/test_dataset_cpu.py), so OpenCV is only touched — and imported — when there is something to undo
        if "distortions" in cam and np.any(np.asarray(cam["distortions"]) != 0):
            cv2 = self._cv2()
            if cv2 is not None:
                K, D = cam["intrinsics"], cam["distortions"]
                orig, alpha = cv2.undistort(orig, K, D), cv2.undistort(alpha, K, D)
            else:
                raise NotImplementedError("lens undistortion needs OpenCV (cv2.undistort, as in the reference): install it or "
                                          "undistort the folder first")
        alpha = alpha / 255.0
        img = alpha * orig + (1.0 - alpha) * np.asarray(bg_color)[None, None, :]
        if "target_size" in self.cfg:
            img, alpha = self._resize(img, self.cfg["target_size"], True), self._resize(alpha, self.cfg["target_size"], False)
        elif self.cfg["resize_img_scale"] != 1.0:
            sc = self.cfg["resize_img_scale"]
            img, alpha = self._resize(img, None, True, scale=sc), self._resize(alpha, None, False, scale=sc)
        return img, alpha, orig_W, orig_H

    def crop_image(self, img, mask, K):
        """reference dataset/train.py:176-200 (random crop around the mask's centre, same np.random call sequence)."""
        crop_w, crop_h = self.cfg["crop_size"]
        h, w, _ = img.shape
        h_center, w_center, _ = np.stack(np.nonzero(mask), axis=-1).mean(axis=0).astype(int)
        h_center = min(h_center, h - (crop_h + 1) // 2) if h_center + (crop_h + 1) // 2 > h else h_center
        h_center = crop_h // 2 if h_center - crop_h // 2 < 0 else h_center
        w_center = min(w_center, w - (crop_w + 1) // 2) if w_center + (crop_w + 1) // 2 > w else w_center
        w_center = crop_w // 2 if w_center - crop_w // 2 < 0 else w_center
        h_left, w_left = h_center - crop_h // 2, w_center - crop_w // 2
        while True:
            rand_w = np.random.randint(max(0, w_left - 50), min(w_left + 50, w - crop_w + 1))
            rand_h = np.random.randint(max(0, h_left - 50), min(h_left + 50, h - crop_h + 1))
            crop_mask = mask[rand_h:rand_h + crop_h, rand_w:rand_w + crop_w]
            if np.sum(crop_mask) < 20:
                continue
            K_new = K.copy()
            K_new[0, 2] -= rand_w
            K_new[1, 2] -= rand_h
            return img[rand_h:rand_h + crop_h, rand_w:rand_w + crop_w], crop_mask, K_new

    def get_total_frames(self):
        return len(self.framelist)

    def __len__(self):
        return len(self.framelist)

    def __getitem__(self, idx):
        name = self.framelist[idx]
        res = {"frame_name": name}
        bgcolor = (np.random.rand(3) * 255.0).astype("float32") if self.bgcolor is None else np.array(self.bgcolor, dtype="float32")
        res["bgcolor"] = bgcolor / 255.0
        if self.prefetch:
            img, alpha, orig_W, orig_H = self.preload[name]
            img = alpha * img + (1.0 - alpha) * bgcolor[None, None, :]
        else:
            img, alpha, orig_W, orig_H = self.load_image(name, bgcolor)
        img = (img / 255.0).astype("float32")
        info = self.mesh_infos[name]
        dst_poses, tpose = info["poses"].astype("float32"), info["tpose_joints"].astype("float32")
        K = self.cameras[name]["intrinsics"][:3, :3].copy()
        if "target_size" in self.cfg:
            scale_w, scale_h = self.cfg["target_size"][0] / orig_W, self.cfg["target_size"][1] / orig_H
        else:
            scale_w, scale_h = self.cfg["resize_img_scale"]
        K[:1] *= scale_w
        K[1:2] *= scale_h
        E, global_tfms = apply_global_tfm_to_camera(self.cameras[name]["extrinsics"], info["Rh"].astype("float32"),
                                                    info["Th"].astype("float32"), return_global_tfms=True)
        res["global_tfms"] = global_tfms
        if list(self.cfg["crop_size"]) != [-1, -1]:
            img, alpha, K = self.crop_image(img, alpha, K)
        res.update({"K": K.astype(np.float32), "E": E.astype(np.float32), "target_rgbs": img,
                    "target_masks": alpha[:, :, 0].astype(np.float32)})
        dst_Rs, dst_Ts = body_pose_to_body_RTs(dst_poses, tpose)
        res.update({"dst_poses": dst_poses, "dst_Rs": dst_Rs, "dst_Ts": dst_Ts, "cnl_gtfms": canonical_global_tfms(self.canonical_joints),
                    "dst_posevec": dst_poses.reshape(-1)[3:] + 1e-2, "joints": get_joints_from_pose(dst_poses, tpose),
                    "dst_tpose_joints": tpose})
        return res

    def get_canonical_info(self):
        """reference dataset/train.py:286-300: what ``Model(model_cfg, canonical_info)`` is built from."""
        return {"canonical_joints": self.canonical_joints,
                "canonical_bbox": {"min_xyz": self.canonical_bbox["min_xyz"], "max_xyz": self.canonical_bbox["max_xyz"],
                                   "scale_xyz": self.canonical_bbox["max_xyz"] - self.canonical_bbox["min_xyz"]},
                "canonical_vertex": self.canonical_vertex, "canonical_lbs_weights": self.canonical_lbs_weights,
                "edges": self.edges, "faces": self.faces}


class NovelViewDataset(Dataset):
    """Mirror of reference ``dataset/test.py::Dataset`` (:28-283) — the ZJU-MoCap novel-view / novel-pose evaluation reader
    of ``eval.py --type view|pose`` (eval.py:214-247) and of train.py's periodic evaluation (train.py:211-220).  Poses and
    the canonical mesh come from the processed folder (``dataset_path``), cameras and pictures from the RAW ZJU-MoCap
    capture (``raw_dataset_path``): ``annots.npy`` (``cams`` K / R / T in millimetres / D per view),
    ``Camera_B<v+1>/<frame:06d>.jpg``, ``mask/…png`` OR-ed with ``mask_cihp/…png``.  Items are ordered frame-major,
    view-minor; every picture is undistorted and halved (LANCZOS4 / LINEAR) like the reference's.  Same constructor
    arguments and item keys.  Pinned by the reference's own reader run with the real OpenCV on a synthetic capture
    (tests/golden/golden_dataset_zju_views.npz)."""

    RESIZE = 0.5                                                        # dataset/test.py:21-24

    def __init__(self, raw_dataset_path, dataset_path, test_type="view", bgcolor=None, exclude_training_view=True,
                 exclude_view=0, skip=30, **_):
        super().__init__(dataset_path, bgcolor=bgcolor)
        self.raw_dataset_path = raw_dataset_path
        annots = np.load(os.path.join(raw_dataset_path, "annots.npy"), allow_pickle=True).item()
        cams = annots["cams"]
        self.cameras = {}
        for view_id in range(len(cams["K"])):
            if exclude_training_view and view_id == exclude_view:
                continue
            E = np.eye(4)
            E[:3, :3] = np.array(cams["R"])[view_id].astype("float32")
            E[:3, 3] = (np.array(cams["T"])[view_id].astype("float32") / 1000.0)[:3, 0]
            self.cameras[view_id] = {"intrinsics": np.array(cams["K"])[view_id].astype("float32"), "extrinsics": E,
                                     "distortions": np.array(cams["D"])[view_id].astype("float32")[:, 0]}
        names = sorted(os.path.splitext(n)[0] for n in os.listdir(self.image_dir) if n.endswith(".png"))
        if test_type == "view":                                          # monohuman's split
            names = names[:-(len(names) // 5)]
        elif test_type == "pose":
            names = names[-(len(names) // 5):]
        else:
            raise NotImplementedError(f"unknown test_type {test_type}")
        self.framelist = names[::skip]

    def __len__(self):
        return len(self.framelist) * len(self.cameras)

    def load_mask(self, rel_png):
        def binary(folder):
            return (_load_rgb(os.path.join(self.raw_dataset_path, folder, rel_png))[:, :, 0] != 0).astype(np.uint8)
        msk = (binary("mask") | binary("mask_cihp")).astype(np.uint8)
        msk[msk == 1] = 255
        return msk

    def load_view_image(self, view_id, frame_id, bg_color):
        cam_dir = f"Camera_B{view_id + 1}"
        orig = _load_rgb(os.path.join(self.raw_dataset_path, cam_dir, f"{frame_id:06d}.jpg"))
        alpha = self.load_mask(os.path.join(cam_dir, f"{frame_id:06d}.png"))
        cam = self.cameras[view_id]
        if np.any(np.asarray(cam["distortions"]) != 0):
            cv2 = self._cv2()
            if cv2 is None:
                raise NotImplementedError("lens undistortion needs OpenCV (cv2.undistort, as in the reference)")
            orig, alpha = cv2.undistort(orig, cam["intrinsics"], cam["distortions"]), cv2.undistort(alpha, cam["intrinsics"], cam["distortions"])
        alpha = (alpha / 255.0)[:, :, None]
        img = alpha * orig + (1.0 - alpha) * np.asarray(bg_color)[None, None, :]
        sc = (self.RESIZE, self.RESIZE)
        img, alpha = self._resize(img, None, True, scale=sc), self._resize(alpha, None, False, scale=sc)
        return img, alpha.reshape(alpha.shape[0], alpha.shape[1])        # cv2.resize drops the single channel

    def __getitem__(self, idx):
        view_id = sorted(self.cameras)[idx % len(self.cameras)]
        name = self.framelist[idx // len(self.cameras)]
        frame_id = int(name.split("_")[1])
        bgcolor = (np.random.rand(3) * 255.0).astype("float32") if self.bgcolor is None else np.array(self.bgcolor, dtype="float32")
        img, alpha = self.load_view_image(view_id, frame_id, bgcolor)
        info = self.mesh_infos[name]
        dst_poses, tpose = info["poses"].astype("float32"), info["tpose_joints"].astype("float32")
        K = self.cameras[view_id]["intrinsics"][:3, :3].copy()
        K[:2] *= self.RESIZE
        E = apply_global_tfm_to_camera(self.cameras[view_id]["extrinsics"], info["Rh"].astype("float32"), info["Th"].astype("float32"))
        dst_Rs, dst_Ts = body_pose_to_body_RTs(dst_poses, tpose)
        return {"frame_name": f"Camera_B{view_id + 1}_{name}", "K": K.astype(np.float32), "E": E.astype(np.float32),
                "target_rgbs": (img / 255.0).astype("float32"), "target_masks": alpha.astype(np.float32),
                "dst_Rs": dst_Rs, "dst_Ts": dst_Ts, "cnl_gtfms": canonical_global_tfms(self.canonical_joints),
                "dst_posevec": dst_poses.reshape(-1)[3:] + 1e-2}


def rotate_camera_by_frame_idx(extrinsics, frame_idx, trans=None, rotate_axis="y", period=196, inv_angle=False):
    """reference utils/camera_util.py:5-109 (``rotate_camera_by_frame_idx`` -> ``_update_extrinsics``): the camera of
    ``extrinsics`` turned by 2 pi frame_idx / period about a world axis through ``trans``."""
    angle = 2 * np.pi * (frame_idx / period)
    if inv_angle:
        angle = -angle
    inv_E = np.linalg.inv(np.asarray(extrinsics))
    camrot, campos = inv_E[:3, :3], inv_E[:3, 3].copy()
    if trans is not None:
        campos = campos - trans
    if camrot.T[1, 1] < 0.0:
        angle = -angle
    vec = np.zeros(3)
    vec[{"x": 0, "y": 1, "z": 2}[rotate_axis]] = angle
    grot = _rodrigues(vec).astype("float32")
    rot_campos, rot_camrot = grot.dot(campos), grot.dot(camrot)
    if trans is not None:
        rot_campos = rot_campos + trans
    E = np.identity(4)
    E[:3, :3] = rot_camrot.T
    E[:3, 3] = -rot_camrot.T.dot(rot_campos)
    return E


class FreeviewDataset(Dataset):
    """Mirror of reference ``dataset/freeview.py::Dataset`` (``eval.py --type freeview``, eval.py:262-277): one training
    frame's pose seen from ``total_frames`` cameras on a circle around the subject.  Same constructor and item keys."""

    ROT_CAM_PARAMS = {"zju_mocap": {"rotate_axis": "z", "inv_angle": True}, "wild": {"rotate_axis": "y", "inv_angle": False}}

    def __init__(self, dataset_path, frame_idx=0, total_frames=100, keyfilter=None, bgcolor=None, src_type="zju_mocap",
                 target_size=None, **_):
        super().__init__(dataset_path, keyfilter=keyfilter, bgcolor=bgcolor if bgcolor is not None else [255.0, 255.0, 255.0],
                         target_size=target_size)
        self.train_frame_idx, self.total_frames, self.src_type = frame_idx, total_frames, src_type
        self.train_frame_name = self.framelist[frame_idx]
        self.train_camera = self.cameras[self.train_frame_name]
        self.train_mesh_info = self.mesh_infos[self.train_frame_name]

    def __len__(self):
        return self.total_frames

    def get_freeview_camera(self, E, frame_idx, total_frames, trans=None):
        E = rotate_camera_by_frame_idx(E, frame_idx, period=total_frames, trans=trans, **self.ROT_CAM_PARAMS[self.src_type])
        return self.train_camera["intrinsics"].copy(), E

    def __getitem__(self, idx):
        img, alpha, orig_W, orig_H = self.load_image(self.train_frame_name, np.array(self.bgcolor, dtype="float32"))
        info = self.train_mesh_info
        dst_poses, tpose = info["poses"].astype("float32").reshape(-1), info["tpose_joints"].astype("float32")
        Rh, Th = info["Rh"].astype("float32"), info["Th"].astype("float32")
        K, E = self.get_freeview_camera(self.train_camera["extrinsics"], idx, self.total_frames, trans=Th)
        if "target_size" in self.cfg:
            scale_w, scale_h = self.cfg["target_size"][0] / orig_W, self.cfg["target_size"][1] / orig_H
        else:
            scale_w, scale_h = self.cfg["resize_img_scale"]
        K[:1] *= scale_w
        K[1:2] *= scale_h
        E = apply_global_tfm_to_camera(E, Rh, Th)
        dst_Rs, dst_Ts = body_pose_to_body_RTs(dst_poses, tpose)
        return {"frame_name": self.train_frame_name + f"_v{idx:04d}", "K": K.astype(np.float32), "E": E.astype(np.float32),
                "target_rgbs": (img / 255.0).astype("float32"), "dst_Rs": dst_Rs, "dst_Ts": dst_Ts,
                "cnl_gtfms": canonical_global_tfms(self.canonical_joints), "dst_posevec": dst_poses[3:] + 1e-2}


def get_camrot(campos, lookat=None, up=None, inv_camera=False):
    """reference utils/camera_util.py:52-83: rows right / up / forward of a camera at ``campos`` looking at ``lookat``."""
    lookat = np.array([0.0, 0.0, 0.0], dtype=np.float32) if lookat is None else lookat
    if up is None:
        up = np.array([0.0, 1.0, 0.0], dtype=np.float32)
        if inv_camera:
            up[1] *= -1.0
    forward = lookat - campos
    forward = forward / np.linalg.norm(forward)
    right = np.cross(up, forward)
    right = right / np.linalg.norm(right)
    up = np.cross(forward, right)
    up = up / np.linalg.norm(up)
    return np.array([right, up, forward], dtype=np.float32)


class NewPoseDataset(Dataset):
    """Mirror of reference ``dataset/newpose.py::Dataset`` (``eval.py --type pose_mdm``, eval.py:248-261): the avatar driven
    by a motion file — an ``.npy`` dictionary with ``thetas_ori`` ``[24,3,N]`` (axis-angle per joint; a torch tensor in the
    files MDM writes, an array is accepted too) and ``root_translation`` ``[3,N]`` — seen from one fixed 512 x 512 camera
    (radius 8, focal 1250, height 1.2).  Targets are zero images, as in the reference.  Unlike the reference it does not
    open (and then discard) ``images/frame_<idx>.png`` for every pose, so the motion may be longer than the training set."""

    RENDER_SIZE = 512
    CAM_PARAMS = {"radius": 8.0, "focal": 1250.0}

    def __init__(self, dataset_path, pose_path, keyfilter=None, bgcolor=(0.0, 0.0, 0.0), debug=False, src_type="wild", **_):
        super().__init__(dataset_path, keyfilter=keyfilter, bgcolor=list(bgcolor) if bgcolor is not None else None)
        self.pose_path, self.src_type = pose_path, src_type
        self.pose_infos = self.load_mdm_pose_infos(pose_path)
        self.total_frames = len(self.pose_infos["Rh"])
        K, E = self.setup_camera(self.RENDER_SIZE, **self.CAM_PARAMS)
        self.camera = {"K": [K] * self.total_frames, "E": [E] * self.total_frames}

    @staticmethod
    def setup_camera(img_size, radius, focal):
        y = 1.2
        campos = np.array([0.0, y, radius], dtype="float32")
        camrot = get_camrot(campos, lookat=np.array([0, y, 0.0]), inv_camera=True)
        E = np.eye(4, dtype="float32")
        E[:3, :3] = camrot
        E[:3, 3] = -camrot.dot(campos)
        K = np.eye(3, dtype="float32")
        K[0, 0] = K[1, 1] = focal
        K[:2, 2] = img_size / 2.0
        return K, E

    @staticmethod
    def load_mdm_pose_infos(path):
        data = dict(np.load(path, allow_pickle=True).item())
        thetas = data["thetas_ori"]
        thetas = thetas.cpu().numpy() if hasattr(thetas, "cpu") else np.asarray(thetas)
        poses = np.transpose(thetas, (2, 0, 1)).copy()
        Rh = poses[:, 0].copy()
        Th = np.transpose(np.asarray(data["root_translation"]), (1, 0))
        poses[:, 0] = 0.0
        return {"poses": poses.reshape(poses.shape[0], -1), "Rh": Rh, "Th": Th}

    def __len__(self):
        return self.total_frames

    def __getitem__(self, idx):
        dst_poses = self.pose_infos["poses"][idx].astype("float32")
        tpose = self.canonical_joints
        Rh, Th = self.pose_infos["Rh"][idx].astype("float32"), self.pose_infos["Th"][0].astype("float32")
        E = apply_global_tfm_to_camera(self.camera["E"][idx], Rh, Th - self.canonical_joints[0])
        dst_Rs, dst_Ts = body_pose_to_body_RTs(dst_poses, tpose)
        H = W = self.RENDER_SIZE
        return {"frame_name": f"frame_{idx:06d}", "dst_poses": dst_poses, "dst_tpose_joints": tpose,
                "K": self.camera["K"][idx].copy().astype(np.float32), "E": E.astype(np.float32),
                "target_rgbs": np.zeros([H, W, 3], dtype=np.float32), "target_masks": np.zeros([H, W], dtype=np.float32),
                "dst_Rs": dst_Rs, "dst_Ts": dst_Ts, "cnl_gtfms": canonical_global_tfms(self.canonical_joints),
                "dst_posevec": dst_poses[3:] + 1e-2, "joints": get_joints_from_pose(dst_poses, tpose)}


# -------------------------------------------------------------------------------------------------------- checkpoints
def save_checkpoint(path, model, optimizer_state=None, n_iter=0):
    """The reference's checkpoint file (train.py:289-294, :372-376): {'iter', 'network': state_dict, 'optimizer'}."""
    import torch
    torch.save({"iter": int(n_iter), "network": model.state_dict(), "optimizer": optimizer_state if optimizer_state is not None else {}}, path)


def model_from_checkpoint(model_cfg, ckpt, map_location="cpu", **model_kwargs):
    """Build ``gomavatar_b200.model.Model`` from a reference checkpoint (a path or the loaded dict) and load it.

    The reference resumes by REPLAYING its mesh subdivisions (``model.subdivide()`` for every ``subdivide_iters`` entry passed,
    train.py:275-279; trimesh-based) so that the parameter shapes match before ``load_state_dict``.  The subdivided topology
    is itself part of the checkpoint — ``faces``, ``lbs_weights`` and ``vertices`` are registered buffers / parameters
    (models/model.py:58-85, :151-170) — so the model is constructed from those tensors directly, whatever number of
    subdivisions produced them.  Returns (model, iteration)."""
    import torch
    from .model import Model
    if isinstance(ckpt, (str, os.PathLike)):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    sd = ckpt["network"] if "network" in ckpt else ckpt
    info = {"canonical_vertex": sd["vertices"].detach().cpu().numpy().T.copy(),
            "canonical_lbs_weights": sd["lbs_weights"].detach().cpu().numpy()[:-1].T.copy(),
            "faces": sd["faces"].detach().cpu().numpy()}
    model = Model(model_cfg, info, **model_kwargs)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match the model built from cfg: missing {list(missing)}, unexpected {list(unexpected)}")
    return model, int(ckpt.get("iter", 0)) if isinstance(ckpt, dict) else 0
