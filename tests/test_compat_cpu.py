"""``gomavatar_b200.compat``: the stand-ins that let the reference's train.py / eval.py run unchanged,
and ``gomavatar_b200.meshes.Meshes``.  Host logic only — no GPU.  The first test needs the reference checkout and is
skipped where it does not exist (the GPU box); the others are self-contained."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_scripts_import_and_run_their_own_code_on_the_stand_ins(golden_dir):
    """tests/host_harness/compat_reference_check.py in a subprocess: the reference's train.py, eval.py and train_pose.py are
    imported byte-unchanged; train.py resolves ``Model`` to this package's; the reference's OWN ``unpack`` + ``compute_loss``
    (through its own mesh_laplacian_smoothing, our Meshes and the pytorch3d.loss stand-in) reproduce golden_loss.npz; the
    reference's OWN ``Model.subdivide`` on the trimesh / Meshes stand-ins reproduces golden_subdivide.npz."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_harness", "compat_reference_check.py"), REF, golden_dir],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["model_class"] == "gomavatar_b200.model.Model" and out["ref_model_class"] == "models.model.Model"
    assert set(out["install"]["shims"]) >= {"pytorch3d", "trimesh"}
    assert set(out["loss"]) == set(out["loss_golden"])
    for k, ref in out["loss_golden"].items():
        assert abs(out["loss"][k] - ref) <= 1e-6 * abs(ref) + 1e-9, (k, out["loss"][k], ref)
    assert abs(out["total"] - out["total_golden"]) <= 1e-6 * out["total_golden"]
    # this package's Model built from the reference's real CfgNode (exps/zju-mocap_377.yaml) == the reference's own Model:
    # every module present, the same state-dict keys and shapes (checkpoints interchange), the same Adam param groups
    # (names, sizes, learning rates after the reference's update_lr) — train.py:262-266, :166-175
    assert out["ours_modules"] == {"pose_refinement_module": "PoseRefinementModule", "non_rigid_module": "NonRigidModule",
                                   "normal_renderer": "Renderer", "shadow_module": "FusedShadowModule"}
    assert out["ours_state"] == out["ref_state"] and len(out["ours_state"]) > 20
    assert out["ours_groups"] == out["ref_groups"] and [g[0] for g in out["ours_groups"]][0] == "lbs_weights"
    assert out["subdivide_iters"] == [50001]
    assert out["conn_equal"] and out["subdivide_conn_equal"] and out["edge_length_close"]
    assert all(out["subdivide_equal"].values()), out["subdivide_equal"]


@pytest.fixture()
def shims():
    import gomavatar_b200.compat as compat
    info = compat.install(b200_model=False)
    yield info
    compat.uninstall()


def test_meshes_accessors_match_dense_restatements():
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.meshes import Meshes
    from gomavatar_b200.model import mesh_edges
    sc = S.make_humanoid(2000, seed=3)
    f = torch.from_numpy(sc.faces).long()
    g = torch.Generator().manual_seed(0)
    v = torch.from_numpy(sc.vertices).double()
    vb = torch.stack([v, v + 0.01 * torch.randn(v.shape, dtype=torch.float64, generator=g)])
    m = Meshes(vb, f[None])
    V, F = v.shape[0], f.shape[0]
    assert len(m) == 2 and not m.isempty() and m.num_verts_per_mesh().tolist() == [V, V]
    assert m.verts_packed().shape == (2 * V, 3) and m.faces_packed().shape == (2 * F, 3)
    assert torch.equal(m.faces_packed()[F:], f + V) and torch.equal(m.verts_packed_to_mesh_idx(), torch.arange(2).repeat_interleave(V))
    # edges: unique (min, max) pairs in ascending order; lengths equal the model's target_edge_length
    e = m.edges_packed()
    E = 3 * F // 2
    assert e.shape == (2 * E, 2) and bool((e[:, 0] < e[:, 1]).all())
    key = e[:, 0] * (2 * V) + e[:, 1]
    assert bool((key[1:] > key[:-1]).all())
    tel, conn = mesh_edges(sc.faces.astype(np.int64), sc.vertices)
    np.testing.assert_allclose((v[e[:E, 0]] - v[e[:E, 1]]).norm(dim=1).numpy(), tel, rtol=1e-6)
    # column k of faces_packed_to_edges_packed is the edge opposite corner k
    fe = m.faces_packed_to_edges_packed()
    fp = m.faces_packed()
    for k, (a, b) in enumerate(((1, 2), (2, 0), (0, 1))):
        pair = torch.sort(torch.stack([fp[:, a], fp[:, b]], 1), dim=1)[0]
        assert torch.equal(e[fe[:, k]], pair)
    # the reference's get_face_connectivity loop (models/model.py:115-125) on top of it reproduces Model.face_connectivity
    fe1 = fe[:F]
    order = torch.argsort(fe1.reshape(-1), stable=True)
    by_edge = (order // 3).reshape(-1, 2)[: int(fe1.max())]
    assert torch.equal(torch.sort(by_edge, dim=1)[0], torch.from_numpy(conn))
    # uniform Laplacian against the dense matrix
    A = torch.zeros(V, V, dtype=torch.float64)
    A[e[:E, 0], e[:E, 1]] = 1
    A[e[:E, 1], e[:E, 0]] = 1
    Ld = A / A.sum(1, keepdim=True) - torch.eye(V, dtype=torch.float64)
    L = m.laplacian_packed().to_dense().double()
    assert float((L[:V, :V] - Ld).abs().max()) < 1e-6 and float(L[:V, V:].abs().max()) == 0 and float((L[V:, V:] - Ld).abs().max()) < 1e-6
    # vertex normals: area-weighted face normals, normalised
    n = torch.cross(vb[1][f[:, 1]] - vb[1][f[:, 0]], vb[1][f[:, 2]] - vb[1][f[:, 0]], dim=1)
    vn = torch.zeros_like(v)
    for k in range(3):
        vn.index_add_(0, f[:, k], n)
    ref = torch.nn.functional.normalize(vn, eps=1e-6, dim=1)
    assert float((m.verts_normals_padded()[1] - ref).abs().max()) < 1e-12
    with pytest.raises(ValueError):
        Meshes(torch.zeros(4, 2), f)


def test_topology_cache_follows_the_face_tensor():
    from gomavatar_b200.meshes import Meshes
    v = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    f = torch.tensor([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    assert Meshes(v[None], f).edges_packed().shape == (6, 2)
    f2 = f[:2].clone()
    assert Meshes(v[None], f2).edges_packed().shape == (5, 2)
    f2[1] = torch.tensor([1, 2, 3])                                    # in-place edit bumps the version: no stale entry
    assert Meshes(v[None], f2).edges_packed().tolist() == [[0, 1], [0, 2], [1, 2], [1, 3], [2, 3]]


def test_pytorch3d_stand_in(shims):
    assert "pytorch3d" in shims["shims"]
    from pytorch3d.loss import mesh_edge_loss, mesh_laplacian_smoothing, mesh_normal_consistency
    from pytorch3d.structures import Meshes
    from pytorch3d.transforms.so3 import so3_exp_map, so3_log_map
    from gomavatar_b200 import regularizers as RG
    from oracle import geometry as G
    w = torch.randn(200, 3, dtype=torch.float64) * torch.logspace(-4, 0.3, 200, dtype=torch.float64)[:, None]
    R = so3_exp_map(w)
    assert float((R - G.so3_exp_map(w)).abs().max()) < 1e-12
    assert float((R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max()) < 1e-6
    big = w[(w.norm(dim=1) > 0.05) & (w.norm(dim=1) < 3.0)]                   # the principal branch
    assert float((so3_log_map(so3_exp_map(big)) - big).abs().max()) < 1e-8
    with pytest.raises(ValueError):
        so3_exp_map(torch.zeros(3))
    v = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=torch.float64, requires_grad=True)
    f = torch.tensor([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    m = Meshes(v[None], f[None])
    nc = mesh_normal_consistency(m)
    assert abs(float(nc) - float(RG.normal_consistency(v, f))) < 1e-14 and float(nc) > 1.0     # a tetrahedron is sharp
    nc.backward()
    assert torch.isfinite(v.grad).all()
    assert abs(float(mesh_laplacian_smoothing(m)) - float(RG.laplacian_smoothing(v, f))) < 1e-14
    e = m.edges_packed()
    assert abs(float(mesh_edge_loss(m, 1.0)) - float((((v[e[:, 0]] - v[e[:, 1]]).norm(dim=1) - 1.0) ** 2).mean())) < 1e-14
    with pytest.raises(ImportError, match="mesh_renderer"):
        import pytorch3d.renderer  # noqa: F401


def test_small_stand_ins(shims):
    import seaborn as sns
    import trimesh
    from skimage.metrics import structural_similarity
    from termcolor import colored
    from torchmetrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure
    from trimesh.remesh import faces_to_edges, grouping
    from gomavatar_b200._lib import GomError
    served = set(shims["shims"])
    if "seaborn" in served:
        hls = np.array(sns.color_palette("hls", 36))
        assert hls.shape == (36, 3) and hls.min() >= 0 and hls.max() <= 1 and len({tuple(c) for c in hls.round(6)}) == 36
        assert np.array(sns.color_palette("tab10", 24)).shape == (24, 3) and np.array(sns.color_palette("coolwarm", 7)).shape == (7, 3)
    if "termcolor" in served:
        assert colored("x", "cyan") == "x"
    if "trimesh" in served:
        f = np.array([[0, 1, 2], [2, 1, 3]])
        mesh = trimesh.Trimesh(np.eye(4)[:, :3] + np.arange(4)[:, None], f, vertex_attributes={"w": np.ones((4, 2))})
        assert mesh.vertices.dtype == np.float64 and mesh.edges.shape == (6, 2) and "w" in mesh.vertex_attributes
        u, inv = grouping.unique_rows(np.sort(faces_to_edges(f), axis=1))
        assert u.tolist() == [0, 2, 1, 4, 5] and len(inv) == 6
        with pytest.raises(NotImplementedError):
            trimesh.Trimesh(np.zeros((4, 3)), f)
    if "torchmetrics" in served:
        g = torch.Generator().manual_seed(0)
        a = torch.rand(1, 3, 40, 48, generator=g)
        b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
        ssim, psnr = StructuralSimilarityIndexMeasure(data_range=1), PeakSignalNoiseRatio(data_range=1)
        assert abs(float(ssim(a, a)) - 1.0) < 1e-6 and 0.0 < float(ssim(a, b)) < 0.999
        assert abs(float(ssim(a, b)) - float(ssim(b, a))) < 1e-6
        assert abs(float(psnr(a, b)) + 10 * np.log10(float(((a - b) ** 2).mean()))) < 1e-4
        from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity
        with pytest.raises(NotImplementedError):
            LearnedPerceptualImagePatchSimilarity(net_type="alex")
    if "skimage" in served:
        img = np.random.default_rng(0).integers(0, 256, (16, 16, 3)) / 255.0
        with pytest.raises(NotImplementedError):
            structural_similarity(img, img, multichannel=True, gaussian_weights=True)
        with pytest.raises(NotImplementedError):
            structural_similarity(img, img * 0.999, multichannel=True)            # not 8-bit levels
        if not torch.cuda.is_available():
            with pytest.raises(GomError):                                          # the metric kernel has no CPU path
                structural_similarity(img, img, multichannel=True)


def test_install_is_idempotent_and_real_packages_win():
    import gomavatar_b200.compat as compat
    a = compat.install(b200_model=False)
    b = compat.install(b200_model=False)
    assert a == b and sys.path.count(compat.SHIM_DIR) == 1 and sys.path[-2:].count(compat.SHIM_DIR) + sys.path[-2:].count(compat.REPO_ROOT) >= 1
    import importlib.util
    assert "numpy" not in a["shims"] and not compat._provided_by_shim("torch")
    for name in a["shims"]:
        assert os.path.abspath(importlib.util.find_spec(name).origin).startswith(compat.SHIM_DIR)
    compat.uninstall()
    assert compat.SHIM_DIR not in sys.path


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_train_and_eval_main_run_unchanged_up_to_the_first_kernel_call(tmp_path):
    """tests/host_harness/compat_train_main_check.py: the reference's own ``train.main`` on the stand-ins, in this GPU-less
    container — config, logger, TensorBoard, its three Dataset objects on a folder dataset_io wrote, this package's Model
    from its cfg node, param groups, Adam, ``iter_0.pt`` — then the first iteration's ``model(...)`` (train.py:317) must
    stop at the first kernel call with GomError: there is no CPU path to fall back to."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_harness", "compat_train_main_check.py"), REF, str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["stopped"] is not None and "no CPU path" in out["stopped"]
    assert "train.py:main" in out["frames"] and out["frames"][-3:] == ["function.py:apply", "skinning.py:forward", "skinning.py:_need_cuda"]
    assert out["files"] == ["checkpoints", "config.yaml", "log.txt", "tb"]
    assert out["ckpt_keys"] == ["iter", "network", "optimizer"] and out["n_param_groups"] == 7     # lbs, app, xyz, scale, so3, pose, shadow
    assert out["reload"] == ["gomavatar_b200.model", 1, 2000]
    # and the reference's own eval.main on a post-subdivision checkpoint: dataset, Model + subdivide replay (eval.py:300-305),
    # load_state_dict, then the first model(...) (eval.py:341) stops at the same place
    assert out["eval_stopped"] is not None and "no CPU path" in out["eval_stopped"]
    assert "eval.py:main" in out["eval_frames"] and out["eval_frames"][-1] == "skinning.py:_need_cuda"
    assert out["eval_dir"] == ["log_view.txt", "view"]
