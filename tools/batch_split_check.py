"""One GPU: the gradient over frames {0,2} and {1,3} (mean loss each), averaged, against the gradient over {0,1,2,3} — the math
of tests/host_harness/dist_grad_check.py without NCCL, with switches to isolate a batch-composition dependence.

    python tools/batch_split_check.py [--no-lpips] [--precision fp32]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomavatar_b200 import synthetic as S  # noqa: E402
from gomavatar_b200.dist import FlatArena  # noqa: E402
from gomavatar_b200.losses import compute_loss  # noqa: E402
from gomavatar_b200.lpips import LPIPS, seeded_random_trunk  # noqa: E402
from gomavatar_b200.model import Model, default_model_cfg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--no-lpips", action="store_true")
ap.add_argument("--precision", default=None)
ap.add_argument("--img", type=int, default=128)
ap.add_argument("--no-ksplit", action="store_true", help="GOM_CONV_KSPLIT=0: batch-invariant convolutions")
ap.add_argument("--conv-impl", default=None)
ap.add_argument("--first-conv-ffma", action="store_true")
ap.add_argument("--torch-first-conv", action="store_true")
a = ap.parse_args()
if a.no_ksplit:
    os.environ["GOM_CONV_KSPLIT"] = "0"
dev = torch.device("cuda:0")
n_faces, img, n_global = 4000, a.img, 4
scene = S.make_humanoid(n_faces, seed=0)
pr = S.make_params(scene, seed=1)
fr = S.make_frames(scene, n_global, img_size=(img, img), seed=7)
rng = np.random.default_rng(3)
tgt = torch.from_numpy(rng.random((n_global, img, img, 3)).astype(np.float32)).to(dev)
tgt_m = torch.from_numpy((rng.random((n_global, img, img)) > 0.5).astype(np.float32)).to(dev)
heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
kw = {} if a.precision is None else {"conv_precision": a.precision}
if a.conv_impl:
    kw["conv_impl"] = a.conv_impl
lp = None if a.no_lpips else LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)], **kw).to(dev)
if lp is not None and a.first_conv_ffma:
    lp.first_conv_tc = False
if lp is not None and a.torch_first_conv:
    lp.own_first_conv = False


def grad(idx):
    m = Model(default_model_cfg(img_size=(img, img)), scene.canonical_info(), strict_raster=False).to(dev).train()
    with torch.no_grad():
        m.so3.copy_(torch.from_numpy(pr["so3"])); m.scale.copy_(torch.from_numpy(pr["scale"]))
        m.appearance_module.appearance.copy_(torch.from_numpy(pr["appearance"]))
    arena = FlatArena(m)
    d = {k: torch.from_numpy(fr[k][idx]).to(dev) for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor")}
    arena.zero_grad()
    rgb, mask, _ = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"], dst_posevec=d["dst_posevec"], bgcolor=d["bgcolor"])
    loss, terms, _ = compute_loss(rgb, mask, d["bgcolor"], tgt[idx], tgt_m[idx], lpips_func=lp)
    loss.backward()
    names = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
    return arena.grad.clone(), arena.slices, names, float(loss), {k: float(v) for k, v in terms.items()}


g02, sl, names, l02, t02 = grad([0, 2])
g13, _, _, l13, t13 = grad([1, 3])
gall, _, _, lall, tall = grad([0, 1, 2, 3])
g = 0.5 * (g02 + g13)
rel = {}
for (name, p), (off, k) in zip(names, sl):
    x, y = g[off:off + k], gall[off:off + k]
    rel[name] = float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
print(json.dumps({"no_ksplit": a.no_ksplit, "env_pair": os.environ.get("GOM_CONV_PAIR"), "no_lpips": a.no_lpips, "precision": a.precision, "conv_impl": a.conv_impl, "ffma": a.first_conv_ffma, "torch_first": a.torch_first_conv, "rel": rel,
                  "loss_split_mean": 0.5 * (l02 + l13), "loss_all": lall, "terms_split": {k: 0.5 * (t02[k] + t13[k]) for k in t02}, "terms_all": tall}))
