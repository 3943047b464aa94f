#!/usr/bin/env python
"""A minimal training loop on the B200-native hot path, shaped like the reference's train.py:309-386 (forward ->
compute_loss -> backward -> optimizer step -> exponential lr decay), on a synthetic SMPL-topology subject because no
dataset is available offline.  A "teacher" parameter set renders the targets; the student starts from perturbed
geometry / reference-init colours and is fitted with L1 + mask + LPIPS.

    python examples/train_synthetic.py --iters 200 --faces 13776 --img 256 --frames 4
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gomavatar_b200 import synthetic as S                      # noqa: E402
from gomavatar_b200.dist import ArenaAdam, FlatArena           # noqa: E402
from gomavatar_b200.losses import compute_loss, unpack        # noqa: E402
from gomavatar_b200.lpips import LPIPS, seeded_random_trunk    # noqa: E402
from gomavatar_b200.metrics import eval_metrics                # noqa: E402
from gomavatar_b200.model import Model, default_model_cfg      # noqa: E402


def make_problem(n_faces, img, n_frames, device, seed=0):
    t = torch.from_numpy
    scene = S.make_humanoid(n_faces, seed=seed)
    model = Model(default_model_cfg(img_size=(img, img)), scene.canonical_info(), strict_raster=False).to(device)
    frames = {k: t(v).to(device) for k, v in S.make_frames(scene, n_frames, img_size=(img, img), seed=seed + 5).items()}
    teacher = S.make_params(scene, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    with torch.no_grad():                                       # targets: the teacher's render over the frame's background
        keep = [p.detach().clone() for p in (model.so3, model.scale, model.appearance_module.appearance)]
        model.so3.copy_(t(teacher["so3"])); model.scale.copy_(t(teacher["scale"]))
        model.appearance_module.appearance.copy_(t(teacher["appearance"]))
        rgb, mask, _ = model(frames["K"], frames["E"], frames["cnl_gtfms"], frames["dst_Rs"], frames["dst_Ts"])
        tgt_rgb = unpack(rgb, mask, frames["bgcolor"]).clamp(0, 1).contiguous()
        tgt_mask = mask.clamp(0, 1).contiguous()
        for p, k in zip((model.so3, model.scale, model.appearance_module.appearance), keep):
            p.copy_(k)                                           # the student restarts from the reference initialisation
        model.vertices.add_(t(rng.normal(0, 2e-3, tuple(model.vertices.shape)).astype(np.float32)).to(device))
    return scene, model, frames, tgt_rgb, tgt_mask


def train(model, frames, tgt_rgb, tgt_mask, iters, lpips=None, lr=5e-3, lr_decay=0.1, decay_steps=None, log=None,
          subdivide_iters=()):
    def make_optimizer():                                         # also after every subdivision (reference train.py:343-346)
        arena = FlatArena(model)
        groups = model.get_param_groups({"lr": {"appearance": lr, "canonical_geometry": lr, "canonical_geometry_xyz": lr * 0.1}})
        opt = ArenaAdam(arena, groups)
        return arena, opt, [g["lr"] for g in opt.param_groups]
    arena, opt, base = make_optimizer()
    history = []
    for it in range(iters):
        arena.zero_grad()
        rgb, mask, _ = model(frames["K"], frames["E"], frames["cnl_gtfms"], frames["dst_Rs"], frames["dst_Ts"],
                             dst_posevec=frames["dst_posevec"], i_iter=it, bgcolor=frames["bgcolor"])
        loss, terms, rgb_u = compute_loss(rgb, mask, frames["bgcolor"], tgt_rgb, tgt_mask, lpips_func=lpips)
        loss.backward()
        opt.step(grad_scale=arena.all_reduce_sum())
        if it in subdivide_iters:                                 # reference train.py:341-346 (cfg.model.subdivide_iters)
            model.subdivide()
            arena, opt, base = make_optimizer()
            if log:
                log(f"iter {it:5d}  subdivided: {model.vertices.shape[1]} vertices, {model.faces.shape[0]} faces")
        if decay_steps:                                           # reference train.py:166-175
            for g, b in zip(opt.param_groups, base):
                g["lr"] = b * lr_decay ** (it / decay_steps)
        if it % max(1, iters // 10) == 0 or it == iters - 1:
            with torch.no_grad():
                m = eval_metrics(rgb_u.clamp(0, 1), tgt_rgb)
            history.append((it, float(loss), float(m["psnr"].mean()), float(m["ssim"].mean())))
            if log:
                log(f"iter {it:5d}  loss {history[-1][1]:.5f}  psnr {history[-1][2]:.2f} dB  ssim {history[-1][3]:.4f}  "
                    + "  ".join(f"{k} {float(v):.5f}" for k, v in terms.items()))
    return history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--faces", type=int, default=13776)
    ap.add_argument("--img", type=int, default=256)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--no-lpips", action="store_true")
    ap.add_argument("--subdivide-iters", type=int, nargs="*", default=[])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = make_problem(a.faces, a.img, a.frames, dev)
    lp = None
    if not a.no_lpips:
        heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
        lp = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)]).to(dev)
    train(model.train(), frames, tgt_rgb, tgt_mask, a.iters, lpips=lp, decay_steps=a.iters, log=print,
          subdivide_iters=tuple(a.subdivide_iters))
