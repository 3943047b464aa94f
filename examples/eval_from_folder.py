#!/usr/bin/env python
"""The reference's eval.py main loop (:183-366, the ``--type train`` / non-ZJU ``--type view`` branch that reads a
processed-dataset folder) on the B200-native path: reference-format folder -> ``Dataset`` -> checkpoint (with the
subdivision replay of eval.py:300-305) -> ``Model`` in eval mode -> ``unpack`` with the clamp of eval.py:80-83 ->
``to_8b_image`` -> PSNR / SSIM / LPIPS x 1000 (``gomavatar_b200.metrics.Evaluator``: one ``gom_eval_metrics`` launch per
batch instead of skimage on the CPU per frame) -> PNGs + ``metric_<type>.npy`` in the reference's format.

    python examples/train_from_folder.py --data /tmp/gom_subject --iters 200
    python examples/eval_from_folder.py  --data /tmp/gom_subject
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gomavatar_b200 import dataset_io as IO                     # noqa: E402
from gomavatar_b200.metrics import Evaluator                    # noqa: E402
from gomavatar_b200.model import Model                          # noqa: E402
from train_from_folder import collate, model_cfg                # noqa: E402


def unpack(rgbs, masks, bgcolors):
    """eval.py:80-83 (the training ``unpack`` + clamp)."""
    rgbs = rgbs * masks.unsqueeze(-1) + bgcolors[:, None, None, :] * (1 - masks).unsqueeze(-1)
    return torch.clamp(rgbs, min=0, max=1)


def load_model(cfg, canonical_info, ckpt_path, n_subdivisions, device):
    """eval.py:297-319: canonical mesh -> ``subdivide(need_face_connectivity=False)`` per passed subdivision ->
    ``load_state_dict(strict=False)`` -> eval mode.  ``eval_mode`` (eval.py:186) skips the soft silhouette."""
    cfg = dict(cfg, eval_mode=True)
    model = Model(cfg, canonical_info)
    for _ in range(n_subdivisions):
        model.subdivide(need_face_connectivity=False)
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    model.load_state_dict(ckpt["network"], strict=False)
    return model.to(device).eval(), int(ckpt.get("iter", 0))


def evaluate(data, ckpt_path, img, device, out_dir, bgcolor=(0.0, 0.0, 0.0), skip=1, batch=4, n_subdivisions=0, lpips=None,
             eval_type="train", log=None):
    ds = IO.Dataset(data, bgcolor=list(bgcolor), skip=skip, target_size=[img, img])           # eval.py:254-260
    loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=False, drop_last=False,
                                         collate_fn=lambda items: (collate(items), [it["frame_name"] for it in items]))
    model, n_iter = load_model(model_cfg(img), ds.get_canonical_info(), ckpt_path, n_subdivisions, device)
    save_dir = os.path.join(out_dir, "eval", eval_type)
    os.makedirs(save_dir, exist_ok=True)
    evaluator = Evaluator(lpips, device=device)
    bg = torch.tensor(bgcolor, dtype=torch.float32, device=device)[None] / 255.0
    n, t0 = 0, time.perf_counter()
    for b, names in loader:
        b = {k: v.to(device, non_blocking=True) for k, v in b.items()}
        with torch.no_grad():
            pred, mask, _ = model(b["K"], b["E"], b["cnl_gtfms"], b["dst_Rs"], b["dst_Ts"], b["dst_posevec"])
            pred = unpack(pred, mask, bg.expand(pred.shape[0], 3))                           # eval.py:345-347
            pred_8b = evaluator.evaluate_batch(pred, b["target_rgbs"], return_8b=True).cpu().numpy()
        for name, im in zip(names, pred_8b):
            Image.fromarray(im).save(os.path.join(save_dir, name + ".png"))                   # eval.py:365
        n += len(names)
    dt = time.perf_counter() - t0
    per_frame = {"mse": list(evaluator.mse), "psnr": list(evaluator.psnr), "ssim": list(evaluator.ssim), "lpips": list(evaluator.lpips)}
    summary = evaluator.summarize(os.path.join(out_dir, "eval", f"metric_{eval_type}.npy"))
    if log:
        log(f"checkpoint iter {n_iter}: {n} frames in {dt:.2f} s ({n / dt:.1f} frames/s incl. PNG writing)  " +
            "  ".join(f"{k} {v:.4f}" for k, v in summary.items()))
    return summary, per_frame, save_dir


def make_dataset(eval_type, data, img, bgcolor, raw=None, skip=1, frame_idx=0, n_frames=100, pose_path=None, exclude_view=0):
    """The dataset selection of eval.py:214-277.  Returns (dataset, has_ground_truth)."""
    if eval_type == "train":
        return IO.Dataset(data, bgcolor=list(bgcolor), skip=skip, target_size=[img, img]), True
    if eval_type == "view":
        if raw is not None:                                                                   # cfg.dataset.test_view.name == 'zju-mocap'
            return IO.NovelViewDataset(raw, data, test_type="view", skip=skip, exclude_view=exclude_view, bgcolor=list(bgcolor)), True
        return IO.Dataset(data, bgcolor=list(bgcolor), skip=skip, target_size=[img, img]), True
    if eval_type == "pose":
        return IO.NovelViewDataset(raw, data, test_type="pose", skip=skip, exclude_training_view=False, bgcolor=list(bgcolor)), True
    if eval_type == "freeview":
        return IO.FreeviewDataset(data, frame_idx, total_frames=n_frames, target_size=[img, img]), False
    if eval_type == "pose_mdm":
        return IO.NewPoseDataset(data, pose_path), False
    raise ValueError(f"unknown evaluation type {eval_type!r}")


RENDER_KEYS = ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "target_rgbs")


def render(eval_type, data, ckpt_path, img, device, out_dir, bgcolor=(0.0, 0.0, 0.0), batch=4, n_subdivisions=0, lpips=None,
           log=None, **dataset_kwargs):
    """eval.py's loop for every ``--type`` (``evaluate`` above is its 'train' branch, kept as tested): metrics only where the
    reader has ground truth (view / pose / train), PNGs always; ``pose`` / ``pose_mdm`` switch the pose refinement off
    (eval.py:326-328) and ``pose_mdm`` renders at 512 x 512 (eval.py:188-190)."""
    if eval_type == "pose_mdm":
        img = 512
    ds, has_gt = make_dataset(eval_type, data, img, bgcolor, **dataset_kwargs)
    item0 = ds[0]
    H, W = item0["target_rgbs"].shape[:2]

    def collate_any(items):
        return ({k: torch.from_numpy(np.stack([np.asarray(it[k], dtype=np.float32) for it in items])) for k in RENDER_KEYS},
                [it["frame_name"] for it in items])
    loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=False, drop_last=False, collate_fn=collate_any)
    cfg = dict(model_cfg(img), img_size=[W, H])
    model, n_iter = load_model(cfg, ds.get_canonical_info(), ckpt_path, n_subdivisions, device)
    if eval_type in ("pose", "pose_mdm"):
        model.pose_refinement_module = None
    save_dir = os.path.join(out_dir, "eval", eval_type)
    os.makedirs(save_dir, exist_ok=True)
    evaluator = Evaluator(lpips, device=device)
    bg = torch.tensor(bgcolor, dtype=torch.float32, device=device)[None] / 255.0
    n = 0
    for b, names in loader:
        b = {k: v.to(device, non_blocking=True) for k, v in b.items()}
        with torch.no_grad():
            pred, mask, _ = model(b["K"], b["E"], b["cnl_gtfms"], b["dst_Rs"], b["dst_Ts"], b["dst_posevec"])
            pred = unpack(pred, mask, bg.expand(pred.shape[0], 3))
            if has_gt:
                pred_8b = evaluator.evaluate_batch(pred, b["target_rgbs"], return_8b=True).cpu().numpy()
            else:
                pred_8b = (255.0 * pred.clamp(0, 1)).to(torch.uint8).cpu().numpy()             # to_8b_image
        for name, im in zip(names, pred_8b):
            Image.fromarray(im).save(os.path.join(save_dir, name + ".png"))
        n += len(names)
    summary = evaluator.summarize(os.path.join(out_dir, "eval", f"metric_{eval_type}.npy")) if has_gt else {}
    if log:
        log(f"checkpoint iter {n_iter}: {n} frames -> {save_dir}  " + "  ".join(f"{k} {v:.4f}" for k, v in summary.items()))
    return summary, save_dir


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="/tmp/gom_subject")
    ap.add_argument("--ckpt", default=None, help="default: the latest iter_*.pt under <data>/checkpoints")
    ap.add_argument("--img", type=int, default=128)
    ap.add_argument("--bgcolor", type=float, default=0.0, help="0..255 like eval.py --bgcolor")
    ap.add_argument("--skip", type=int, default=1)
    ap.add_argument("--subdivisions", type=int, default=0, help="len(cfg.model.subdivide_iters) of the run that made the checkpoint")
    ap.add_argument("--type", default="train", choices=["train", "view", "pose", "freeview", "pose_mdm"], help="eval.py --type")
    ap.add_argument("--raw", default=None, help="raw ZJU-MoCap capture (annots.npy, Camera_B*/) for --type view / pose")
    ap.add_argument("--frame_idx", type=int, default=0, help="freeview only")
    ap.add_argument("--n_frames", type=int, default=100, help="freeview only")
    ap.add_argument("--pose_path", default=None, help="pose_mdm only: MDM-format motion file")
    a = ap.parse_args()
    ck = a.ckpt
    if ck is None:                                                                            # eval.py:308-312
        d = os.path.join(a.data, "checkpoints")
        ck = os.path.join(d, "iter_%d.pt" % max(int(f.split("_")[-1][:-3]) for f in os.listdir(d) if "pose" not in f))
    if a.type == "train":
        evaluate(a.data, ck, a.img, torch.device("cuda:0"), a.data, bgcolor=(a.bgcolor,) * 3, skip=a.skip,
                 n_subdivisions=a.subdivisions, log=print)
    else:
        render(a.type, a.data, ck, a.img, torch.device("cuda:0"), a.data, bgcolor=(a.bgcolor,) * 3, n_subdivisions=a.subdivisions,
               log=print, raw=a.raw, skip=a.skip, frame_idx=a.frame_idx, n_frames=a.n_frames, pose_path=a.pose_path)
