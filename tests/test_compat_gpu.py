"""GPU: the ``gomavatar_b200.compat`` stand-ins on device tensors — what the reference's unchanged train.py / eval.py
would call with this package's ``Model`` outputs."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


@pytest.fixture()
def shims():
    import gomavatar_b200.compat as compat
    info = compat.install(b200_model=False)
    yield info
    compat.uninstall()


def test_model_outputs_feed_the_reference_style_mesh_losses(shims):
    """outputs['mesh'] / ['mesh_canonical'] through pytorch3d.loss (stand-in) and the reference's Laplacian recipe
    (utils/network_util.py:748-792: L = laplacian_packed(); (L @ verts).norm ** 2 mean) equal the fused kernels' terms."""
    import train_synthetic as TS
    from pytorch3d.loss import mesh_normal_consistency
    from gomavatar_b200 import regularizers as RG
    dev = torch.device("cuda:0")
    scene, model, frames, tgt_rgb, tgt_mask = TS.make_problem(2000, 64, 1, dev)
    model.strict_raster = True
    model.train()
    rgb, mask, out = model(frames["K"], frames["E"], frames["cnl_gtfms"], frames["dst_Rs"], frames["dst_Ts"],
                           dst_posevec=frames["dst_posevec"], i_iter=0, bgcolor=frames["bgcolor"])
    mesh, canon = out["mesh"], out["mesh_canonical"]
    assert len(mesh) == 1 and mesh.verts_packed().shape == (model.vertices.shape[1], 3) and mesh.device.type == "cuda"
    assert torch.equal(canon.verts_packed(), model.vertices.t())
    topo = RG.mesh_topology(model.faces, model.face_connectivity, model.vertices.shape[1])
    lap_k, nc_k, cc_k = RG.fused_mesh_regularizers(out["vertices_observation"], out["colors"], topo)
    L = mesh.laplacian_packed()
    lap_ref = (L.mm(mesh.verts_packed()).norm(dim=1) ** 2).mean()
    nc_ref = mesh_normal_consistency(mesh)
    assert abs(float(lap_k) - float(lap_ref)) <= 1e-4 * float(lap_ref) + 1e-9, (float(lap_k), float(lap_ref))
    assert abs(float(nc_k) - float(nc_ref)) <= 1e-4 * float(nc_ref) + 1e-9, (float(nc_k), float(nc_ref))
    (10.0 * lap_ref + 0.1 * nc_ref).backward()                     # gradients reach the parameters through LBS
    assert torch.isfinite(model.vertices.grad).all() and float(model.vertices.grad.abs().sum()) > 0


def test_skimage_stand_in_is_the_metric_kernel(shims):
    if "skimage" not in shims["shims"]:
        pytest.skip("a real scikit-image is installed")
    from skimage.metrics import structural_similarity
    from gomavatar_b200.metrics import eval_metrics
    from oracle import losses as OL
    rng = np.random.default_rng(3)
    gt = rng.integers(0, 256, (48, 40, 3)).astype(np.float64) / 255.0
    pred = np.clip(np.rint((gt + rng.normal(0, 0.05, gt.shape)) * 255.0), 0, 255) / 255.0
    got = structural_similarity(pred, gt, multichannel=True)                      # eval.py:107
    dev = torch.device("cuda:0")
    m = eval_metrics(torch.from_numpy(pred).float().to(dev), torch.from_numpy(gt).float().to(dev), quantize=False)
    assert abs(got - float(m["ssim"][0])) < 1e-12                 # same kernel; fp64 atomics make the last bit order-dependent
    assert abs(got - float(OL.ssim(pred, gt))) < 1e-9


def test_unchanged_reference_train_and_eval_run_on_the_gpu(tmp_path):
    """SURVEY.md §8 f-4, end to end: the reference's own train.main (20 iterations incl. a mesh subdivision, the periodic
    evaluate(), checkpoints), ``--resume`` and eval.main, all byte-unchanged, on ``gomavatar_b200.compat``
    (tests/host_harness/compat_train_gpu_run.py).  Needs a reference checkout next to the GPU: /root/reference, or a copy in the
    git-ignored ``_ref_scratch/`` of the snapshot (the tree is never committed); skipped otherwise.  Log of the run this was
    developed with: profiles/r5_compat_unchanged_train_eval_gpu.log."""
    import json
    import subprocess
    ref = next((p for p in ("/root/reference", os.path.join(ROOT, "_ref_scratch")) if os.path.exists(os.path.join(p, "train.py"))), None)
    if ref is None:
        pytest.skip("no reference checkout on this machine")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_harness", "compat_train_gpu_run.py"), ref, str(tmp_path)],
                         capture_output=True, text=True, timeout=900)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
    assert res.returncode == 0 and lines, res.stdout[-3000:] + res.stderr[-3000:]
    out = json.loads(lines[-1][7:])
    assert out["model_module"] == "gomavatar_b200.model"
    assert out["checkpoints"] == ["iter_0.pt", "iter_10.pt", "iter_20.pt"] and "iter_20.pt" in out["checkpoints_after_resume"]
    assert out["ckpt_faces_after_subdivision"] == 4 * 4000 and out["ckpt_optimizer_groups"] >= 5
    assert any("evaluate on test" in ln for ln in out["log_lines"]) and any("subdivide at iter 8" in ln for ln in out["log_lines"])
    assert "view" in out["eval_dir"]
