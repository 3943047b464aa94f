"""GPU: the reference's OWN ``train.main`` and ``eval.main`` (train.py:178-390, eval.py:183-366, byte-unchanged) executed on
``gomavatar_b200.compat`` for a few iterations — datasets, config, TensorBoard, Model.forward on libgom_b200.so, the
reference's compute_loss, torch.optim.Adam, a mesh subdivision in the middle, checkpoints, the periodic evaluate(), a
``--resume`` and finally eval.py on the last checkpoint.  argv: reference root, scratch dir.  SURVEY.md §8 f-4.

The reference tree does not exist on the GPU box: it travels there in a git-ignored scratch directory of the snapshot and is
never committed.  One concession, not part of the product: ``LPIPS(net='vgg')`` downloads torchvision's
ImageNet weights (no network), so ``train.LPIPS`` / ``eval.LPIPS`` resolve to this package's LPIPS (row a-12's drop-in: same
call signature, in-tree v0.1 heads from the reference's utils/lpips/weights) over a seeded random VGG16 trunk."""
import argparse
import json
import os
import re
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ref, scratch = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])

import gomavatar_b200.compat as compat  # noqa: E402
from gomavatar_b200 import dataset_io as IO  # noqa: E402
from gomavatar_b200 import synthetic as S  # noqa: E402

served = compat.install(ref)
data, n, img = os.path.join(scratch, "subject"), 8, 128
scene = S.make_humanoid(4000, seed=0)
cams = [S.make_camera(azimuth=2 * np.pi * i / n, img_size=(img, img), focal=537.0 * img / 512, base_size=img) for i in range(n)]
# targets: renders of a perturbed teacher, so that the losses have something to fit
from gomavatar_b200.model import Model, default_model_cfg  # noqa: E402
dev = torch.device("cuda:0")
poses = S.make_poses(n, seed=5)
rng = np.random.default_rng(0)
imgs = (rng.random((n, img, img, 3)) * 40 + 60).astype(np.uint8)
masks = np.zeros((n, img, img), np.uint8)
masks[:, img // 5: -img // 5, img // 3: -img // 3] = 255
IO.write_synthetic_dataset(data, scene, poses, cams, imgs, masks)
total = 20
yaml_path = os.path.join(scratch, "explore.yaml")
with open(yaml_path, "w") as f:
    f.write(f"""
exp_name: "explore"
save_dir: "{os.path.join(scratch, 'log')}"
random_bgcolor: true
bgcolor: [0., 0., 0.]
img_size: [{img}, {img}]
dataset:
  train: {{dataset_path: "{data}", batch_size: 1, num_workers: 0}}
  test_view: {{name: "synthetic", dataset_path: "{data}", batch_size: 1, num_workers: 0, skip: 4}}
  test_on_train: {{batch_size: 1, num_workers: 0}}
model:
  img_size: [{img}, {img}]
  subdivide_iters: [8]
  canonical_geometry: {{deform_scale: true, deform_so3: true}}
  non_rigid: {{name: 'basic', condition_code_size: 69, mlp_width: 128, mlp_depth: 6, skips: [4], multires: 6, i_embed: 0, kick_in_iter: 12, full_band_iter: 16}}
  pose_refinement: {{name: 'mlp', embedding_size: 69, mlp_width: 256, mlp_depth: 4, kick_in_iter: 4}}
  normal_renderer: {{name: 'mesh', soft_mask: true, sigma: 0.00001}}
  shadow_module: {{name: 'basic', mlp_width: 128, mlp_depth: 3, skips: [4], multires: 6, i_embed: 0}}
train:
  total_iters: {total}
  log_freq: 2
  save_freq: 10
  eval_freq: 10
""")
os.chdir(ref)
import train  # noqa: E402  (reference, unchanged)
from gomavatar_b200.lpips import LPIPS as B200LPIPS, load_head_weights, seeded_random_trunk  # noqa: E402


def lpips_factory(net="vgg", **_):
    return B200LPIPS(seeded_random_trunk(0), load_head_weights(os.path.join(ref, "utils", "lpips", "weights", "v0.1", "vgg.pth")))


import eval as ref_eval  # noqa: E402  (reference, unchanged; train.evaluate builds eval.Evaluator, eval.py:93)
train.LPIPS = lpips_factory
ref_eval.LPIPS = lpips_factory
out = {"shims": served, "model_module": train.Model.__module__, "gpu": torch.cuda.get_device_name(0)}
t0 = time.time()
train.main(argparse.Namespace(cfg=yaml_path, resume=False))
out["train_seconds"] = time.time() - t0
log_dir = os.path.join(scratch, "log", "explore")
out["files"] = sorted(os.listdir(log_dir))
out["checkpoints"] = sorted(os.listdir(os.path.join(log_dir, "checkpoints")))
text = ""
for fn in os.listdir(log_dir):
    if fn.endswith(".log") or fn.endswith(".txt"):
        text += open(os.path.join(log_dir, fn)).read()
out["log_lines"] = [ln for ln in text.splitlines() if "iter " in ln and "loss" in ln][:40] + \
    [ln for ln in text.splitlines() if "evaluate on" in ln or "subdivide" in ln or "saved to" in ln]
ck = torch.load(os.path.join(log_dir, "checkpoints", f"iter_{total}.pt"), map_location="cpu", weights_only=False)
out["ckpt_faces_after_subdivision"] = int(ck["network"]["faces"].shape[0])
out["ckpt_optimizer_groups"] = len(ck["optimizer"]["param_groups"])
# ---- resume (train.py:269-286: replays the subdivision, loads network + optimizer) for 4 more iterations
cfg_text = open(yaml_path).read().replace(f"total_iters: {total}", f"total_iters: {total + 4}")
open(yaml_path, "w").write(cfg_text)
train.main(argparse.Namespace(cfg=yaml_path, resume=True))
out["checkpoints_after_resume"] = sorted(os.listdir(os.path.join(log_dir, "checkpoints")))
# ---- eval.py on the latest checkpoint
ref_eval.main(argparse.Namespace(cfg=yaml_path, type="view", iter=None, frame_idx=0, n_frames=1, bgcolor=None, pose_path=None))
ev = os.path.join(log_dir, "eval")
out["eval_dir"] = sorted(os.listdir(ev))
for root_, _, files in os.walk(ev):
    for fn in files:
        if fn.endswith(".npy") or fn.endswith(".txt"):
            out.setdefault("eval_files", []).append(os.path.relpath(os.path.join(root_, fn), ev))
print("RESULT " + json.dumps(out))
