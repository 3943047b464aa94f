"""Executed in a subprocess by tests/test_compat_cpu.py (the reference's top-level packages are called ``utils``,
``models``, ``train``, ``eval`` ... and must not leak into the pytest process).  argv: reference root, golden dir.
Runs the reference's OWN train.py / eval.py / models/model.py code, imported byte-unchanged, on top of
``gomavatar_b200.compat`` and prints one JSON line."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ref, golden = sys.argv[1], sys.argv[2]
out = {}

import gomavatar_b200.compat as compat  # noqa: E402

out["install"] = compat.install(ref)
os.chdir(ref)
import train  # noqa: E402  (reference, unchanged)
import eval as ref_eval  # noqa: E402
import train_pose  # noqa: E402,F401

out["model_class"] = f"{train.Model.__module__}.{train.Model.__name__}"
out["evaluator"] = ref_eval.Evaluator.__name__

# ---- the reference's own compute_loss + unpack on OUR Meshes objects and pytorch3d.loss stand-in (CPU tensors)
from gomavatar_b200.meshes import Meshes  # noqa: E402
from utils import lpips as ref_lpips  # noqa: E402

g = np.load(os.path.join(golden, "golden_loss.npz"))
t = lambda k: torch.from_numpy(g[k])
torch.manual_seed(0)
lp = ref_lpips.LPIPS(net="vgg", pnet_rand=True, verbose=False)
NS = types.SimpleNamespace
cfgl = NS(rgb=NS(coeff=1.0), mask=NS(coeff=5.0), lpips=NS(coeff=1.0), laplacian=NS(coeff_canonical=0.0, coeff_observation=10.0),
          normal=NS(mask_dilate=True, kernel_size=7, coeff_mask=1.0, coeff_consist=0.10), color_consist=NS(coeff=0.050))
mesh = Meshes(t("verts")[None], t("faces"))
outputs = {"mesh": mesh, "mesh_canonical": mesh, "normal_mask": t("normal_mask"), "colors": t("colors"),
           "face_connectivity": t("face_connectivity")}
with torch.no_grad():
    rgb = train.unpack(t("rgb_raw"), t("mask_pred"), t("bgcolor"))
    total, losses = train.compute_loss(rgb, t("mask_pred"), outputs, t("rgb_gt"), t("mask_gt"), cfgl, None, 0, lpips_func=lp)
out["loss"] = {k: float(v["unscaled"]) for k, v in losses.items()}
out["loss_golden"] = {k[len("unscaled."):]: float(g[k]) for k in g.files if k.startswith("unscaled.")}
out["total"], out["total_golden"] = float(total), float(g["total"])

# ---- this package's Model built from the reference's REAL config node (yacs CfgNode of exps/zju-mocap_377.yaml) + the
#      reference's own optimizer recipe (train.py:264-266) and update_lr (train.py:166-175) on its param groups
from configs import make_cfg  # noqa: E402

s = np.load(os.path.join(golden, "golden_subdivide.npz"))
info = {"faces": s["in.faces"], "canonical_vertex": s["in.vertices"].T.copy(), "canonical_lbs_weights": s["in.lbs_weights"][:-1].T.copy(),
        "edges": np.zeros((0, 2), np.int64)}
cfg = make_cfg("exps/zju-mocap_377.yaml")
ours = train.Model(cfg.model, info)
out["ours_modules"] = {k: type(getattr(ours, k)).__name__ for k in ("pose_refinement_module", "non_rigid_module", "normal_renderer", "shadow_module")}
out["ours_state"] = {k: list(v.shape) for k, v in ours.state_dict().items()}
opt = torch.optim.Adam(ours.get_param_groups(cfg.train), betas=(0.9, 0.999))
train.update_lr(opt, 1000, cfg.train)
out["ours_groups"] = [[g["name"], float(g["lr"]), sum(int(p.numel()) for p in g["params"])] for g in opt.param_groups]
out["subdivide_iters"] = list(cfg.model.subdivide_iters)

# ---- the reference's own Model (models/model.py, unchanged) + its own subdivide() on the trimesh / Meshes stand-ins
compat.install(ref, b200_model=False)
sys.modules.pop("models.model", None)
torch.Tensor.cuda = lambda self, *a, **k: self              # model.py:58,60 call .cuda() in the constructor
import models.model as ref_model  # noqa: E402

# full ZJU config except the PyTorch3D mesh renderer (no parameters of its own): state-dict keys / shapes, param groups
cfg_full = make_cfg("exps/zju-mocap_377.yaml")
cfg_full.model.normal_renderer.name = "none"
import contextlib, io  # noqa: E402
with contextlib.redirect_stdout(io.StringIO()):                 # get_param_groups prints every tensor (model.py:322)
    ref_full = ref_model.Model(cfg_full.model, info)
    ropt = torch.optim.Adam(ref_full.get_param_groups(cfg_full.train), betas=(0.9, 0.999))
train.update_lr(ropt, 1000, cfg_full.train)
out["ref_state"] = {k: list(v.shape) for k, v in ref_full.state_dict().items()}
out["ref_groups"] = [[g["name"], float(g["lr"]), sum(int(p.numel()) for p in g["params"])] for g in ropt.param_groups]

cfg = make_cfg("exps/zju-mocap_377.yaml")
cfg.model.img_size = [64, 64]
for node in ("normal_renderer", "shadow_module", "non_rigid", "pose_refinement"):
    getattr(cfg.model, node).name = "none"
if "eval_mode" not in cfg.model:
    cfg.model.eval_mode = False
m = ref_model.Model(cfg.model, info)
out["ref_model_class"] = f"{type(m).__module__}.{type(m).__name__}"
out["conn_equal"] = bool(np.array_equal(m.face_connectivity.numpy(), s["in.face_connectivity"]))
with torch.no_grad():
    m.so3.copy_(torch.from_numpy(s["in.so3"]))
    m.scale.copy_(torch.from_numpy(s["in.scale"]))
    m.appearance_module.appearance.copy_(torch.from_numpy(s["in.appearance_module.appearance"]))
m.subdivide()
sd = m.state_dict()
out["subdivide_equal"] = {k: bool(np.array_equal(sd[k].detach().numpy(), s[f"s1.{k}"]))
                          for k in ("vertices", "faces", "lbs_weights", "so3", "scale", "appearance_module.appearance")}
out["subdivide_conn_equal"] = bool(np.array_equal(m.face_connectivity.numpy(), s["s1.face_connectivity"]))
out["edge_length_close"] = bool(np.allclose(sd["target_edge_length"].numpy(), s["s1.target_edge_length"], rtol=1e-6))
print("RESULT " + json.dumps(out))
