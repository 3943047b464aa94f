"""Executed in a subprocess by tests/test_compat_cpu.py.  argv: reference root, scratch dir.
The reference's OWN ``train.main`` (train.py:178-390, byte-unchanged) on ``gomavatar_b200.compat``, in a container without
a GPU: config, logging, TensorBoard, the three dataset objects reading a folder ``dataset_io`` wrote, this package's
``Model`` built from the reference's config node, its param groups, Adam and the ``iter_0.pt`` checkpoint all run; the first
iteration then reaches the first kernel call and must stop THERE with ``GomError`` — the product has no CPU path.
Two concessions to this container, neither part of the product: ``.cuda()`` is a no-op (no device), and ``train.LPIPS`` is
a stub (``LPIPS(net='vgg')`` downloads torchvision weights: no network)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ref, scratch = sys.argv[1], sys.argv[2]

import gomavatar_b200.compat as compat  # noqa: E402
from gomavatar_b200 import dataset_io as IO  # noqa: E402
from gomavatar_b200 import synthetic as S  # noqa: E402
from gomavatar_b200._lib import GomError  # noqa: E402

compat.install(ref)
data, n, img = os.path.join(scratch, "subject"), 4, 64
scene = S.make_humanoid(2000, seed=0)
cams = [S.make_camera(azimuth=2 * np.pi * i / n, img_size=(img, img), focal=537.0 * img / 512, base_size=img) for i in range(n)]
IO.write_synthetic_dataset(data, scene, S.make_poses(n, seed=5), cams, np.full((n, img, img, 3), 90, np.uint8), np.full((n, img, img), 255, np.uint8))
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
  test_view: {{name: "synthetic", dataset_path: "{data}", batch_size: 1, num_workers: 0, skip: 2}}
  test_on_train: {{batch_size: 1, num_workers: 0}}
model:
  img_size: [{img}, {img}]
  subdivide_iters: [3]
  canonical_geometry: {{deform_scale: true, deform_so3: true}}
  pose_refinement: {{name: 'mlp', embedding_size: 69, mlp_width: 256, mlp_depth: 4, kick_in_iter: 0}}
  normal_renderer: {{name: 'mesh', soft_mask: true, sigma: 0.00001}}
  shadow_module: {{name: 'basic', mlp_width: 128, mlp_depth: 3, skips: [4], multires: 6, i_embed: 0}}
train:
  total_iters: 5
""")
os.chdir(ref)
import train  # noqa: E402  (reference, unchanged)

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self


class _NoLpips(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, a, b):
        return (a - b).abs().mean().reshape(1, 1, 1, 1)


train.LPIPS = _NoLpips
out = {"stopped": None}
try:
    train.main(argparse.Namespace(cfg=yaml_path, resume=False))
except GomError as e:
    import traceback
    tb = traceback.extract_tb(e.__traceback__)
    out["stopped"] = str(e)
    out["frames"] = [f"{os.path.basename(fr.filename)}:{fr.name}" for fr in tb]
log = os.path.join(scratch, "log", "explore")
out["files"] = sorted(os.listdir(log))
ck = torch.load(os.path.join(log, "checkpoints", "iter_0.pt"), map_location="cpu", weights_only=False)
out["ckpt_keys"] = sorted(ck)
out["n_param_groups"] = len(ck["optimizer"]["param_groups"])
model, it = IO.model_from_checkpoint(train.make_cfg(yaml_path).model, ck)
out["reload"] = [type(model).__module__, it, int(model.faces.shape[0])]

# ---- the reference's own eval.main (eval.py:183-366, unchanged) on a checkpoint taken AFTER the subdivision: dataset,
#      Model + subdivide(need_face_connectivity=False) replay (eval.py:300-305), load_state_dict(strict=False), then the
#      first model(...) call (eval.py:341) must stop at the first kernel
model.subdivide()
IO.save_checkpoint(os.path.join(log, "checkpoints", "iter_5.pt"), model, n_iter=5)
import eval as ref_eval  # noqa: E402

ref_eval.LPIPS = _NoLpips
out["eval_stopped"] = None
try:
    ref_eval.main(argparse.Namespace(cfg=yaml_path, type="view", iter=None, frame_idx=0, n_frames=1, bgcolor=None, pose_path=None))
except GomError as e:
    import traceback
    out["eval_stopped"] = str(e)
    out["eval_frames"] = [f"{os.path.basename(fr.filename)}:{fr.name}" for fr in traceback.extract_tb(e.__traceback__)]
out["eval_dir"] = sorted(os.listdir(os.path.join(log, "eval")))
print("RESULT " + json.dumps(out))
