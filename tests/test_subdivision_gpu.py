"""GPU: the reference's mid-training subdivision (train.py:341-346, models/model.py:136-179) on the B200 path.  The host
logic itself is pinned on the CPU (tests/test_subdivision_cpu.py); here: every kernel accepts the subdivided model, the
render is the same surface, the regulariser kernels rebuild their cached topology, and training carries on."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _forward(model, frames, it=0):
    return model(frames["K"], frames["E"], frames["cnl_gtfms"], frames["dst_Rs"], frames["dst_Ts"],
                 dst_posevec=frames["dst_posevec"], i_iter=it, bgcolor=frames["bgcolor"])


def test_subdivided_model_renders_the_same_surface_and_backpropagates():
    import train_synthetic as TS
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = TS.make_problem(2000, 128, 2, dev)
    model.strict_raster = True
    model.train()
    with torch.no_grad():
        rgb0, mask0, _ = _forward(model, frames)
    V0, F0 = model.vertices.shape[1], model.faces.shape[0]
    model.subdivide()
    assert (model.vertices.shape[1], model.faces.shape[0]) == (V0 + 3 * F0 // 2, 4 * F0)
    assert all(t.device.type == "cuda" for t in (model.vertices, model.faces, model.lbs_weights, model.so3, model.scale,
                                                 model.appearance_module.appearance, model.target_edge_length,
                                                 model.face_connectivity))
    rgb1, mask1, out = _forward(model, frames)
    assert out["colors"].shape == (4 * F0, 3) and out["radii"].shape[-1] == 4 * F0
    a, b = mask0 > 0.5, mask1 > 0.5
    iou = float((a & b).sum()) / float((a | b).sum())
    assert iou > 0.85, iou                                    # 4 children tile their parent: the silhouette stays
    assert float((rgb0 - rgb1.detach()).abs().mean()) < 0.05
    loss = (rgb1 - tgt_rgb).abs().mean() + (mask1 - tgt_mask).abs().mean()
    loss.backward()
    for p in (model.vertices, model.so3, model.scale, model.appearance_module.appearance):
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all()
        assert float(p.grad.abs().sum()) > 0


def test_regulariser_kernels_follow_the_new_topology():
    """``regularizers.compute_loss`` caches the mesh topology per connectivity tensor: after ``subdivide()`` the fused
    kernels must equal the torch definitions on the NEW mesh."""
    import train_synthetic as TS
    from types import SimpleNamespace as NS
    from gomavatar_b200 import regularizers as RG
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = TS.make_problem(2000, 64, 2, dev)
    model.strict_raster = True
    model.train()
    cfgl = NS(rgb=NS(coeff=1.0), mask=NS(coeff=5.0), lpips=NS(coeff=0.0),
              laplacian=NS(coeff_canonical=0.0, coeff_observation=10.0),
              normal=NS(mask_dilate=True, kernel_size=7, coeff_mask=0.0, coeff_consist=0.1), color_consist=NS(coeff=0.05))
    for step in range(2):
        rgb, mask, out = _forward(model, frames)
        total, losses = RG.compute_loss(rgb, mask, frames["bgcolor"], tgt_rgb, tgt_mask, out, model, cfgl, lpips_func=None)
        vo = out["vertices_observation"].permute(0, 2, 1).double()
        with torch.no_grad():
            lap = RG.laplacian_smoothing(vo, model.faces)
            nc = RG.normal_consistency(vo, model.faces)
            cc = RG.color_consistency(out["colors"].double(), model.face_connectivity)
        for name, ref in (("laplacian_observation", lap), ("normal_consist", nc), ("color_consist", cc)):
            got = float(losses[name]["unscaled"])
            assert abs(got - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-7, (step, name, got, float(ref))
        total.backward()
        assert torch.isfinite(model.vertices.grad).all()
        if step == 0:
            model.subdivide()


def test_training_continues_across_a_subdivision():
    import train_synthetic as TS
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = TS.make_problem(2000, 64, 4, dev)
    hist = TS.train(model.train(), frames, tgt_rgb, tgt_mask, 60, lpips=None, decay_steps=60, subdivide_iters=(20,))
    assert model.faces.shape[0] == 8000 and model.vertices.shape[1] == 1034 + 3000
    (_, loss0, psnr0, _), (_, loss1, psnr1, _) = hist[0], hist[-1]
    assert loss1 < 0.8 * loss0 and psnr1 > psnr0 + 1.0, hist
    assert int(model.last_raster_aux["status"].max()) == 0
    for p in model.parameters():
        assert torch.isfinite(p).all()


def test_full_model_training_with_a_subdivision_then_resume(tmp_path):
    """examples/train_from_folder.py with ``subdivide_iters``: the mesh normal renderer, the tcgen05 shadow MLP and the
    regulariser kernels run on the subdivided mesh; a checkpoint taken AFTER the subdivision resumes."""
    import train_from_folder as TF
    dev = torch.device("cuda:0")
    data = str(tmp_path / "subject")
    torch.manual_seed(0)
    TF.write_synthetic_subject(data, 2000, 64, 4, dev)
    ck = os.path.join(data, "checkpoints")
    model, hist = TF.train(data, 24, 64, dev, ckpt_dir=ck, save_freq=24, lr=5e-3, subdivide_iters=(12,))
    assert model.faces.shape[0] == 8000 and len(hist) == 24 and all(np.isfinite(hist))
    assert np.mean(hist[-6:]) < 1.25 * np.mean(hist[6:12]), hist                           # no blow-up across the event
    assert os.listdir(ck) == ["iter_24.pt"]
    model2, hist2 = TF.train(data, 30, 64, dev, ckpt_dir=ck, save_freq=0, lr=5e-3, subdivide_iters=(12,))
    assert model2.faces.shape[0] == 8000 and len(hist2) == 6 and all(np.isfinite(hist2))
    assert np.mean(hist2) < 1.25 * np.mean(hist[-6:]), (hist2, hist[-6:])
