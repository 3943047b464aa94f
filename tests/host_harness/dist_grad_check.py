"""Run under ``python -m torch.distributed.run --nproc-per-node N`` (one rank per GPU, NCCL): the frame-sharded step of
gomavatar_b200/dist.py with the REAL model and kernels.  Each rank renders its own frames {rank, rank + N, ...} of one global
batch, takes the mean loss over them, back-propagates into the flat gradient arena and joins the ONE summing all-reduce;
scaled by 1 / N that must equal the gradient a single rank computes over ALL frames with the mean loss (SURVEY.md §8e:
"N-rank semantics = the 1-rank step over the same N x B frames").  Rank 0 computes that reference gradient itself and prints
one JSON line with the relative error per parameter; the Adam step that follows must leave all replicas bit-identical."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

# Batch-invariant convolutions: the 1-rank side pushes 2 x 4 images through the VGG trunk per launch, a rank of the 2-rank side
# 2 x 2; with K-splits allowed the small deep layers pick different splits for the two and the tensor core's fp32 accumulation
# order moves activations at the 1e-5 level (ReLU masks near zero flip: ~1e-3 of the gradient).  Without K-splits every tile
# shape is bit-identical (tools/conv_auto_vs_pinned.py), so the comparison below tests the sharding / all-reduce semantics alone.
os.environ["GOM_CONV_KSPLIT"] = "0"

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gomavatar_b200 import synthetic as S  # noqa: E402
from gomavatar_b200.dist import ArenaAdam, FlatArena, init_from_env, shard_frames  # noqa: E402
from gomavatar_b200.losses import compute_loss  # noqa: E402
from gomavatar_b200.lpips import LPIPS, seeded_random_trunk  # noqa: E402
from gomavatar_b200.model import Model, default_model_cfg  # noqa: E402


def main():
    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n_faces, img, b_local = 4000, 128, 2
    n_global = b_local * world
    scene = S.make_humanoid(n_faces, seed=0)
    pr = S.make_params(scene, seed=1)
    fr = S.make_frames(scene, n_global, img_size=(img, img), seed=7)
    rng = np.random.default_rng(3)
    tgt = torch.from_numpy(rng.random((n_global, img, img, 3)).astype(np.float32)).to(dev)
    tgt_m = torch.from_numpy((rng.random((n_global, img, img)) > 0.5).astype(np.float32)).to(dev)
    heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
    lp = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)]).to(dev)

    def build():
        m = Model(default_model_cfg(img_size=(img, img)), scene.canonical_info(), strict_raster=False).to(dev).train()
        with torch.no_grad():
            m.so3.copy_(torch.from_numpy(pr["so3"])); m.scale.copy_(torch.from_numpy(pr["scale"]))
            m.appearance_module.appearance.copy_(torch.from_numpy(pr["appearance"]))
        return m

    def grad_over(model, arena, idx):
        d = {k: torch.from_numpy(fr[k][idx]).to(dev) for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor")}
        arena.zero_grad()
        rgb, mask, _ = model(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"], dst_posevec=d["dst_posevec"], bgcolor=d["bgcolor"])
        loss, _, _ = compute_loss(rgb, mask, d["bgcolor"], tgt[idx], tgt_m[idx], lpips_func=lp)      # mean over the frames given
        loss.backward()
        return float(loss.detach())

    model = build()
    arena = FlatArena(model)
    arena.broadcast_params()
    groups = model.get_param_groups({"lr": {"appearance": 5e-4, "canonical_geometry": 5e-4, "canonical_geometry_xyz": 5e-4}})
    opt = ArenaAdam(arena, groups)
    mine = shard_frames(n_global, rank, world)
    loss_local = grad_over(model, arena, mine)
    scale = arena.all_reduce_sum()                       # THE collective of the step
    g_dist = (arena.grad * scale).clone()
    opt.step(grad_scale=scale)
    # replicas must stay bit-identical after the step
    mx, mn = arena.data.clone(), arena.data.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    replicas_identical = bool(torch.equal(mx, mn))
    out = None
    if rank == 0:
        ref_model = build()
        ref_arena = FlatArena(ref_model)
        loss_ref = grad_over(ref_model, ref_arena, list(range(n_global)))
        rel = {}
        for (name, p), (off, k) in zip(((n, p) for n, p in ref_model.named_parameters() if p.requires_grad), ref_arena.slices):
            a, b = g_dist[off:off + k], ref_arena.grad[off:off + k]
            rel[name] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        out = {"world": world, "frames_global": n_global, "loss_rank0_local": loss_local, "loss_all_frames": loss_ref,
               "grad_rel_err_vs_1rank": rel, "replicas_identical_after_adam": replicas_identical,
               "nccl": ".".join(str(v) for v in torch.cuda.nccl.version())}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        ok = replicas_identical and max(out["grad_rel_err_vs_1rank"].values()) < 1e-4
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
