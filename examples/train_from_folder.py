#!/usr/bin/env python
"""The reference's train.py loop (:248-386) on the B200-native path, reading a processed-dataset FOLDER in the reference's
format: dataset -> DataLoader -> Model(cfg.model, dataset.get_canonical_info()) -> forward -> compute_loss (L1, mask,
LPIPS, Laplacian / normal / colour regularisers) -> backward -> Adam with exponential lr decay -> checkpoints in the
reference's format, resumable.  No dataset ships offline, so by default a synthetic subject is first WRITTEN to the folder in
that format (teacher renders as images) — point --data at a folder prepared by the reference's scripts/prepare_* instead.

    python examples/train_from_folder.py --data /tmp/gom_subject --iters 200 --img 128 --faces 2000
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gomavatar_b200 import dataset_io as IO                     # noqa: E402
from gomavatar_b200 import synthetic as S                       # noqa: E402
from gomavatar_b200.dist import ArenaAdam, FlatArena            # noqa: E402
from gomavatar_b200.model import Model                          # noqa: E402
from gomavatar_b200.regularizers import compute_loss            # noqa: E402

LOSSES = {"rgb": {"coeff": 1.0}, "mask": {"coeff": 5.0}, "lpips": {"coeff": 1.0},          # configs/default.yaml + exps/zju-mocap_377.yaml
          "laplacian": {"coeff_canonical": 0.0, "coeff_observation": 10.0},
          "normal": {"mask_dilate": True, "kernel_size": 7, "coeff_mask": 1.0, "coeff_consist": 0.10},
          "color_consist": {"coeff": 0.050}}


def model_cfg(img):
    return {"img_size": [img, img], "eval_mode": False,
            "canonical_geometry": {"sigma": 1e-3, "radius_scale": 1.0, "deform_scale": True, "deform_so3": True},
            "appearance": {"color_init": 0.5},
            "normal_renderer": {"name": "mesh", "soft_mask": True, "sigma": 1e-5},
            "shadow_module": {"name": "basic", "mlp_width": 128, "mlp_depth": 3, "skips": [4], "multires": 6, "i_embed": 0}}


def write_synthetic_subject(path, n_faces, img, n_frames, device, seed=0):
    """A teacher avatar rendered from n_frames cameras / poses, stored as a reference-format folder."""
    t = torch.from_numpy
    scene = S.make_humanoid(n_faces, seed=seed)
    teacher = Model(model_cfg(img), scene.canonical_info()).to(device).eval()
    pr = S.make_params(scene, seed=seed + 1)
    with torch.no_grad():
        teacher.so3.copy_(t(pr["so3"])); teacher.scale.copy_(t(pr["scale"])); teacher.appearance_module.appearance.copy_(t(pr["appearance"]))
        teacher.shadow_module.block_mlps[-1].weight.normal_(0, 0.1)
    poses = S.make_poses(n_frames, seed=seed + 5)
    cams = [S.make_camera(azimuth=2 * np.pi * i / n_frames, img_size=(img, img), focal=537.0 * img / 512, base_size=img) for i in range(n_frames)]
    images, masks = [], []
    for i, (K, E) in enumerate(cams):
        Rs, Ts = S.body_pose_to_body_RTs(poses[i], scene.joints)
        with torch.no_grad():
            rgb, mask, _ = teacher(t(K)[None].float().to(device), t(E)[None].float().to(device), t(scene.cnl_gtfms)[None].to(device),
                                   t(Rs)[None].to(device), t(Ts)[None].to(device))
        images.append((rgb[0].clamp(0, 1) * 255).round().byte().cpu().numpy())
        masks.append((mask[0].clamp(0, 1) * 255).round().byte().cpu().numpy())
    IO.write_synthetic_dataset(path, scene, poses, cams, np.stack(images), np.stack(masks))


def collate(items):
    keys = ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor", "target_rgbs", "target_masks")
    return {k: torch.from_numpy(np.stack([np.asarray(it[k], dtype=np.float32) for it in items])) for k in keys}


def train(data, iters, img, device, batch=2, lr=5e-4, ckpt_dir=None, save_freq=0, lpips=None, log=None, subdivide_iters=()):
    ds = IO.Dataset(data, target_size=[img, img])
    loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=True, drop_last=True, collate_fn=collate,
                                         generator=torch.Generator().manual_seed(0))
    n_iter, model, opt_state = 0, None, None
    if ckpt_dir and os.path.isdir(ckpt_dir) and os.listdir(ckpt_dir):                     # train.py:269-286 (--resume)
        last = max(int(f.split("_")[-1][:-3]) for f in os.listdir(ckpt_dir))
        ckpt = torch.load(os.path.join(ckpt_dir, f"iter_{last}.pt"), map_location="cpu", weights_only=False)
        model, n_iter = IO.model_from_checkpoint(model_cfg(img), ckpt, strict_raster=False)
        opt_state = ckpt.get("optimizer") or None                                          # train.py:281
    if model is None:
        model = Model(model_cfg(img), ds.get_canonical_info(), strict_raster=False)
    model = model.to(device).train()
    def make_optimizer():
        arena = FlatArena(model)
        groups = model.get_param_groups({"lr": {"appearance": lr, "canonical_geometry": lr, "canonical_geometry_xyz": lr, "shadow": lr}})
        opt = ArenaAdam(arena, groups)
        return arena, opt, [g["lr"] for g in opt.param_groups]
    arena, opt, base = make_optimizer()
    if opt_state is not None:                  # Adam moments and per-parameter step counts continue where they stopped
        opt.load_state_dict(opt_state)
        base = [g["lr"] / 0.1 ** (n_iter / 100000) for g in opt.param_groups]
    history = []
    while n_iter < iters:
        for b in loader:
            if n_iter >= iters:
                break
            b = {k: v.to(device, non_blocking=True) for k, v in b.items()}
            arena.zero_grad()
            rgbs, masks, outputs = model(b["K"], b["E"], b["cnl_gtfms"], b["dst_Rs"], b["dst_Ts"], dst_posevec=b["dst_posevec"],
                                         i_iter=n_iter, bgcolor=b["bgcolor"])
            loss, terms = compute_loss(rgbs, masks, b["bgcolor"], b["target_rgbs"], b["target_masks"], outputs, model, LOSSES, lpips_func=lpips)
            loss.backward()
            opt.step(grad_scale=arena.all_reduce_sum(), active=model.active_param_groups(n_iter))
            n_iter += 1
            if n_iter in subdivide_iters:                                                  # train.py:341-346: every face -> 4,
                model.subdivide()                                                          # new Parameters -> new optimizer
                arena, opt, base = make_optimizer()
                if log:
                    log(f"iter {n_iter:6d}  subdivided: {model.vertices.shape[1]} vertices, {model.faces.shape[0]} faces")
            for g, b0 in zip(opt.param_groups, base):                                      # train.py:166-175
                g["lr"] = b0 * 0.1 ** (n_iter / 100000)
            history.append(float(loss.detach()))
            aux = model.last_raster_aux                                                    # None right after a subdivision
            if aux is not None and n_iter % max(1, iters // 10) == 0:                      # one sync per log interval
                st = int(aux["status"].max().item())
                if st:     # k_emit dropped (Gaussian, tile) instances: the tile lists of those frames were truncated
                    raise RuntimeError(f"rasterizer status {st} at iteration {n_iter}: instance capacity "
                                       f"{aux['inst_capacity']} exceeded; rebuild the Model with a larger raster_capacity")
            if log and (n_iter % max(1, iters // 10) == 0):
                log(f"iter {n_iter:6d}  loss {history[-1]:.5f}  " + "  ".join(f"{k} {float(v['scaled']):.5f}" for k, v in terms.items()))
            if ckpt_dir and save_freq and n_iter % save_freq == 0:
                os.makedirs(ckpt_dir, exist_ok=True)
                IO.save_checkpoint(os.path.join(ckpt_dir, f"iter_{n_iter}.pt"), model, optimizer_state=opt.state_dict(), n_iter=n_iter)
    return model, history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="/tmp/gom_subject")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--faces", type=int, default=2000)
    ap.add_argument("--img", type=int, default=128)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--subdivide-iters", type=int, nargs="*", default=[])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if not os.path.exists(os.path.join(a.data, "cameras.pkl")):
        write_synthetic_subject(a.data, a.faces, a.img, a.frames, dev)
        print(f"wrote a synthetic subject in the reference's format to {a.data}")
    train(a.data, a.iters, a.img, dev, ckpt_dir=os.path.join(a.data, "checkpoints"), save_freq=max(1, a.iters // 2), log=print,
          subdivide_iters=tuple(a.subdivide_iters))
