"""CPU: on-disk compatibility (gomavatar_b200/dataset_io.py).

tests/golden/golden_dataset.npz holds what the reference's OWN reader (dataset/train.py::Dataset, run by
oracle/make_golden.py::dataset_golden) returned for the folder that ``write_synthetic_dataset`` wrote from the seeded
fixture of oracle/dataset_fixture.py.  Here the same folder is written again and read with ``dataset_io.Dataset``: every item
field, the canonical info and the random-crop branch must agree; and the items must feed ``Model`` (shapes / dtypes)."""
import os
import pickle

import numpy as np
import pytest

from gomavatar_b200 import dataset_io as IO
from oracle import dataset_fixture as DF


@pytest.fixture(scope="module")
def folder(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("dataset"))
    fixture = DF.build(path)
    return path, fixture


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_dataset.npz"))


def test_folder_has_the_reference_layout(folder):
    path, (scene, poses, cams, images, masks) = folder
    assert sorted(os.listdir(path)) == ["avg_betas.npy", "cameras.pkl", "canonical_joints.pkl", "images", "masks", "mesh_infos.pkl"]
    cams_pkl = pickle.load(open(os.path.join(path, "cameras.pkl"), "rb"))
    mesh_pkl = pickle.load(open(os.path.join(path, "mesh_infos.pkl"), "rb"))
    cj = pickle.load(open(os.path.join(path, "canonical_joints.pkl"), "rb"))
    names = sorted(cams_pkl)
    assert names == sorted(mesh_pkl) == [f"frame_{i:06d}" for i in range(DF.N_FRAMES)]
    assert set(cams_pkl[names[0]]) == {"intrinsics", "extrinsics", "distortions"}          # prepare_dataset.py:143-147
    assert set(mesh_pkl[names[0]]) == {"Rh", "Th", "poses", "joints", "tpose_joints"}      # :152-158
    assert set(cj) == {"vertex", "joints", "weights", "edges", "faces"}                    # :188-196
    assert cj["weights"].shape == (scene.n_vertices, 24) and mesh_pkl[names[0]]["poses"].shape == (72,)


def test_items_equal_the_reference_reader(folder, gold):
    path, _ = folder
    ds = IO.Dataset(path, bgcolor=[255.0, 128.0, 0.0], target_size=[DF.W, DF.H])
    assert len(ds) == DF.N_FRAMES
    for i in range(len(ds)):
        item = ds[i]
        keys = {k[len(f"item{i}."):] for k in gold.files if k.startswith(f"item{i}.")}
        assert set(item) == keys, (set(item) ^ keys)
        assert item["frame_name"] == str(gold[f"item{i}.frame_name"])
        for k in keys - {"frame_name"}:
            ref, got = gold[f"item{i}.{k}"], np.asarray(item[k])
            assert got.shape == ref.shape and got.dtype == ref.dtype, (k, got.shape, ref.shape, got.dtype, ref.dtype)
            tol = 0 if k in ("target_rgbs", "target_masks", "bgcolor", "dst_poses", "dst_tpose_joints", "dst_posevec") else 2e-6
            assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (k, np.abs(got - ref).max())
    # the model's inputs come out in the shapes train.py batches (then [1, ...] after the DataLoader)
    item = ds[0]
    assert item["K"].shape == (3, 3) and item["E"].shape == (4, 4) and item["cnl_gtfms"].shape == (24, 4, 4)
    assert item["dst_Rs"].shape == (24, 3, 3) and item["dst_Ts"].shape == (24, 3) and item["dst_posevec"].shape == (69,)
    assert item["target_rgbs"].shape == (DF.H, DF.W, 3) and item["target_masks"].shape == (DF.H, DF.W)


def test_stored_extrinsics_round_trip_through_the_global_transform(folder):
    path, (scene, poses, cams, images, masks) = folder
    ds = IO.Dataset(path, bgcolor=[0.0, 0.0, 0.0], target_size=[DF.W, DF.H])
    for i in range(len(ds)):                              # the writer stored E G; the reader's E inv(G) gives the model's E back
        np.testing.assert_allclose(ds[i]["E"], cams[i][1], atol=2e-6)
        np.testing.assert_allclose(ds[i]["K"], cams[i][0], atol=1e-6)


def test_canonical_info_and_crop_branch_equal_the_reference_reader(folder, gold):
    path, (scene, *_rest) = folder
    ds = IO.Dataset(path, bgcolor=[0.0, 0.0, 0.0], target_size=[DF.W, DF.H], crop_size=[32, 24])
    info = ds.get_canonical_info()
    for k in gold.files:
        if not k.startswith("info."):
            continue
        parts = k.split(".")[1:]
        got = info[parts[0]] if len(parts) == 1 else info[parts[0]][parts[1]]
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), gold[k].astype(np.float64), atol=1e-7)
    np.random.seed(3)                                     # same seed as the golden run: same crop window
    item = ds[1]
    for k in ("K", "target_rgbs", "target_masks"):
        assert item[k].shape == gold[f"crop.{k}"].shape
        assert np.abs(np.asarray(item[k], dtype=np.float64) - gold[f"crop.{k}"]).max() <= 1e-6, k
    # the canonical info builds the model (CPU construction only; the kernels need a GPU)
    from gomavatar_b200.model import Model, default_model_cfg
    m = Model(default_model_cfg(img_size=(DF.W, DF.H)), info)
    assert m.vertices.shape == (3, scene.n_vertices) and m.lbs_weights.shape == (25, scene.n_vertices)


def test_checkpoint_round_trip_in_the_reference_format_including_a_subdivided_mesh(tmp_path):
    """{'iter', 'network', 'optimizer'} like train.py:289-294; a checkpoint taken after a subdivision (4x the faces) is
    restored from its own faces / lbs_weights / vertices tensors, no subdivision replay."""
    import torch
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.model import Model
    cfg = {"img_size": [64, 64], "canonical_geometry": {"sigma": 1e-3, "radius_scale": 1.0, "deform_scale": True, "deform_so3": True},
           "appearance": {"color_init": 0.5},
           "shadow_module": {"name": "basic", "mlp_width": 128, "mlp_depth": 3, "skips": [4], "multires": 6},
           "normal_renderer": {"name": "mesh", "soft_mask": True, "sigma": 1e-5}}
    for n_faces in (2000, 8000):                           # 8000 stands for "after one subdivision of the 2000-face mesh"
        scene = S.make_humanoid(n_faces, seed=0)
        torch.manual_seed(n_faces)
        m = Model(cfg, scene.canonical_info())
        with torch.no_grad():
            m.so3.normal_(0, 0.1)
            m.appearance_module.appearance.uniform_(0, 1)
        path = str(tmp_path / f"iter_{n_faces}.pt")
        IO.save_checkpoint(path, m, optimizer_state={"state": {}, "param_groups": []}, n_iter=1234)
        raw = torch.load(path, weights_only=False)
        assert set(raw) == {"iter", "network", "optimizer"} and raw["iter"] == 1234
        m2, it = IO.model_from_checkpoint(cfg, path)
        assert it == 1234 and m2.faces.shape[0] == n_faces
        sd, sd2 = m.state_dict(), m2.state_dict()
        assert set(sd) == set(sd2)
        for k in sd:
            assert torch.equal(sd[k], sd2[k]), k


def test_opencv_pixel_path_equals_the_reference_reader(tmp_path, golden_dir):
    """tests/golden/golden_dataset_cv2.npz: the reference's own reader with the REAL OpenCV on the fixture folder with lens
    distortion (``cv2.undistort``), a ``target_size`` that differs from the files' (``cv2.resize`` LANCZOS4 / LINEAR) and the
    ``resize_img_scale`` branch.  With OpenCV installed ``dataset_io.Dataset`` makes the same calls: the pixels agree."""
    cv2 = pytest.importorskip("cv2")
    gold = np.load(os.path.join(golden_dir, "golden_dataset_cv2.npz"))
    path = str(tmp_path / "distorted")
    DF.build(path)
    DF.add_distortion(path)
    same_build = str(gold["cv2_version"]) == cv2.__version__
    for tag, kw in (("resized", dict(target_size=[64, 56])), ("halved", dict())):
        ds = IO.Dataset(path, bgcolor=[255.0, 128.0, 0.0], **kw)
        for i in range(len(ds)):
            item = ds[i]
            for k in ("K", "E", "target_rgbs", "target_masks", "bgcolor"):
                ref, got = gold[f"{tag}.item{i}.{k}"], np.asarray(item[k])
                assert got.shape == ref.shape and got.dtype == ref.dtype, (tag, i, k, got.shape, ref.shape, got.dtype, ref.dtype)
                tol = 2e-6 if k in ("K", "E") else (0.0 if same_build else 1e-5)      # pixels: the same library calls
                assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (tag, i, k, np.abs(got - ref).max())
    # distortion did change the picture (the test is not vacuous): frame 1 differs from the undistorted read of the same file
    clean = str(tmp_path / "clean")
    DF.build(clean)
    a = IO.Dataset(clean, bgcolor=[255.0, 128.0, 0.0], target_size=[64, 56])[1]["target_rgbs"]
    b = IO.Dataset(path, bgcolor=[255.0, 128.0, 0.0], target_size=[64, 56])[1]["target_rgbs"]
    assert np.abs(a - b).max() > 0.05


def test_identity_operations_never_touch_opencv(tmp_path, monkeypatch):
    """cv2.undistort with zero coefficients and cv2.resize to the same size return their input bit for bit, so the reader
    skips them (and never imports OpenCV) for folders that need neither — every synthetic folder of the GPU tests."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(20):
        H, W = int(rng.integers(20, 200)), int(rng.integers(20, 200))
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        K = np.array([[rng.uniform(50, 2000), 0, rng.uniform(0, W)], [0, rng.uniform(50, 2000), rng.uniform(0, H)], [0, 0, 1]])
        assert np.array_equal(cv2.undistort(img, K, np.zeros(5)), img)
        f = rng.random((H, W, 3))
        assert np.array_equal(cv2.resize(f, [W, H], interpolation=cv2.INTER_LANCZOS4), f)
        assert np.array_equal(cv2.resize(f, [W, H], interpolation=cv2.INTER_LINEAR), f)
    path = str(tmp_path / "clean")
    DF.build(path)

    def boom():
        raise AssertionError("OpenCV must not be needed here")
    monkeypatch.setattr(IO.Dataset, "_cv2", staticmethod(boom))
    IO.Dataset(path, bgcolor=[0.0, 0.0, 0.0], target_size=[DF.W, DF.H])[0]


def test_without_opencv_distortion_is_refused_and_pillow_resamples(tmp_path, monkeypatch):
    path = str(tmp_path / "distorted")
    DF.build(path)
    monkeypatch.setattr(IO.Dataset, "_cv2", staticmethod(lambda: None))
    ds = IO.Dataset(path, bgcolor=[0.0, 0.0, 0.0], target_size=[64, 56])             # zero distortion: fine without OpenCV
    item = ds[0]
    assert item["target_rgbs"].shape == (56, 64, 3) and item["target_masks"].shape == (56, 64)
    assert np.isfinite(item["target_rgbs"]).all() and -0.5 < item["target_rgbs"].min() and item["target_rgbs"].max() < 1.5   # Lanczos rings on noise
    DF.add_distortion(path)
    ds = IO.Dataset(path, bgcolor=[0.0, 0.0, 0.0], target_size=[64, 56])
    with pytest.raises(NotImplementedError, match="OpenCV"):
        ds[1]


def test_novel_view_reader_equals_the_reference_test_reader(tmp_path, golden_dir):
    """tests/golden/golden_dataset_zju_views.npz: the reference's own ``dataset/test.py::Dataset`` (eval.py --type view /
    pose on ZJU-MoCap, real OpenCV) on a synthetic RAW capture — annots.npy cameras in millimetres with lens distortion,
    jpg pictures, mask OR mask_cihp, the monohuman frame splits, frame-major / view-minor order, the excluded training view."""
    cv2 = pytest.importorskip("cv2")
    gold = np.load(os.path.join(golden_dir, "golden_dataset_zju_views.npz"))
    raw, proc = str(tmp_path / "raw"), str(tmp_path / "processed")
    os.makedirs(raw)
    DF.build_raw_zju(raw, proc)
    same_build = str(gold["cv2_version"]) == cv2.__version__
    for tag, kw in (("view", dict(test_type="view", skip=3, exclude_view=0)),
                    ("pose", dict(test_type="pose", skip=1, exclude_training_view=False))):
        ds = IO.NovelViewDataset(raw, proc, bgcolor=[10.0, 200.0, 90.0], **kw)
        assert len(ds) == int(gold[f"{tag}.len"]) > 0
        for i in range(len(ds)):
            item = ds[i]
            keys = {k[len(f"{tag}.item{i}."):] for k in gold.files if k.startswith(f"{tag}.item{i}.")}
            assert set(item) == keys, set(item) ^ keys
            assert item["frame_name"] == str(gold[f"{tag}.item{i}.frame_name"])
            for k in keys - {"frame_name"}:
                ref, got = gold[f"{tag}.item{i}.{k}"], np.asarray(item[k])
                assert got.shape == ref.shape and got.dtype == ref.dtype, (tag, i, k, got.shape, ref.shape, got.dtype, ref.dtype)
                tol = (0.0 if same_build else 1e-5) if k in ("target_rgbs", "target_masks", "dst_posevec") else 2e-6
                assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (tag, i, k, np.abs(got - ref).max())
    assert np.array_equal(ds.get_canonical_info()["canonical_vertex"], gold["info.canonical_vertex"])
    with pytest.raises(NotImplementedError):
        IO.NovelViewDataset(raw, proc, test_type="tpose")


def test_freeview_reader_equals_the_reference_freeview_reader(folder, golden_dir):
    """tests/golden/golden_dataset_freeview.npz: the reference's own ``dataset/freeview.py::Dataset`` (eval.py --type
    freeview) on the fixture folder — the camera circle (``rotate_camera_by_frame_idx``) in both conventions."""
    pytest.importorskip("cv2")                       # the 'wild' case resamples by 0.5: OpenCV's LANCZOS4, like the reference
    gold = np.load(os.path.join(golden_dir, "golden_dataset_freeview.npz"))
    path, _ = folder
    for tag, kw in (("zju", dict(src_type="zju_mocap", target_size=[DF.W, DF.H])), ("wild", dict(src_type="wild", bgcolor=[0.0, 64.0, 255.0]))):
        ds = IO.FreeviewDataset(path, 1, total_frames=7, **kw)
        assert len(ds) == int(gold[f"{tag}.len"]) == 7
        Es = []
        for i in range(len(ds)):
            item = ds[i]
            keys = {k[len(f"{tag}.item{i}."):] for k in gold.files if k.startswith(f"{tag}.item{i}.")}
            assert set(item) == keys, set(item) ^ keys
            assert item["frame_name"] == str(gold[f"{tag}.item{i}.frame_name"])
            for k in keys - {"frame_name"}:
                ref, got = gold[f"{tag}.item{i}.{k}"], np.asarray(item[k])
                assert got.shape == ref.shape and got.dtype == ref.dtype, (tag, i, k)
                tol = 1e-6 if k == "target_rgbs" else 3e-6
                assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (tag, i, k, np.abs(got - ref).max())
            Es.append(item["E"])
        assert np.abs(Es[0] - Es[3]).max() > 0.1                          # the camera does move


def test_new_pose_reader_equals_the_reference_newpose_reader(folder, tmp_path, golden_dir):
    """tests/golden/golden_dataset_newpose.npz: the reference's own ``dataset/newpose.py::Dataset`` (eval.py --type
    pose_mdm) on the fixture folder and a seeded MDM-format motion file: the fixed camera, the root handling, every key."""
    gold = np.load(os.path.join(golden_dir, "golden_dataset_newpose.npz"))
    path, _ = folder
    motion = DF.write_mdm_motion(str(tmp_path / "motion.npy"), DF.N_FRAMES)
    ds = IO.NewPoseDataset(path, motion)
    assert len(ds) == int(gold["len"]) == DF.N_FRAMES
    for i in range(len(ds)):
        item = ds[i]
        keys = {k[len(f"item{i}."):].replace(".shape", "") for k in gold.files if k.startswith(f"item{i}.")}
        assert set(item) == keys, set(item) ^ keys
        assert item["frame_name"] == str(gold[f"item{i}.frame_name"])
        for k in ("target_rgbs", "target_masks"):
            assert list(item[k].shape) == list(gold[f"item{i}.{k}.shape"]) and item[k].dtype == np.float32 and not item[k].any()
        for k in keys - {"frame_name", "target_rgbs", "target_masks"}:
            ref, got = gold[f"item{i}.{k}"], np.asarray(item[k])
            assert got.shape == ref.shape and got.dtype == ref.dtype, (i, k, got.shape, ref.shape, got.dtype, ref.dtype)
            assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= 3e-6, (i, k, np.abs(got - ref).max())
    # a plain-array motion file (no torch tensor inside) and a motion longer than the training set both work here
    longer = DF.write_mdm_motion(str(tmp_path / "longer.npy"), DF.N_FRAMES + 4, as_torch=False)
    assert len(IO.NewPoseDataset(path, longer)) == DF.N_FRAMES + 4 and IO.NewPoseDataset(path, longer)[DF.N_FRAMES + 3]["K"][0, 0] == 1250.0


def test_eval_example_selects_every_reader_and_reaches_the_kernels(folder, tmp_path):
    """examples/eval_from_folder.py::render — eval.py's ``--type`` switch (train / view / pose / freeview / pose_mdm): the
    reader, the model at the reader's image size, the checkpoint with its subdivision replay, pose refinement off for
    unseen poses; on this GPU-less machine each run must then stop at the first kernel call (no CPU path)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import eval_from_folder as EF
    from gomavatar_b200._lib import GomError
    from gomavatar_b200.model import Model
    path, _ = folder
    raw, proc = str(tmp_path / "raw"), str(tmp_path / "processed")
    os.makedirs(raw)
    DF.build_raw_zju(raw, proc)
    motion = DF.write_mdm_motion(str(tmp_path / "motion.npy"), 5, as_torch=False)
    ds, gt = EF.make_dataset("view", proc, 64, (0.0, 0.0, 0.0), raw=raw, skip=3)
    assert isinstance(ds, IO.NovelViewDataset) and gt and ds[0]["target_rgbs"].shape == (DF.RAW_H // 2, DF.RAW_W // 2, 3)
    assert isinstance(EF.make_dataset("freeview", path, 64, (0.0,) * 3, frame_idx=1, n_frames=5)[0], IO.FreeviewDataset)
    assert isinstance(EF.make_dataset("pose_mdm", path, 64, (0.0,) * 3, pose_path=motion)[0], IO.NewPoseDataset)
    with pytest.raises(ValueError):
        EF.make_dataset("tpose", path, 64, (0.0,) * 3)
    if torch.cuda.is_available():
        pytest.skip("the rest checks the no-CPU-path behaviour")
    # a checkpoint of a once-subdivided model, in the reference's format
    m = Model(EF.model_cfg(64), IO.Dataset(proc).get_canonical_info())
    m.subdivide()
    ck = str(tmp_path / "iter_9.pt")
    IO.save_checkpoint(ck, m, n_iter=9)
    dev = torch.device("cpu")
    for eval_type, kw in (("view", dict(raw=raw, skip=3)), ("pose", dict(raw=raw, skip=1)), ("freeview", dict(frame_idx=1, n_frames=3)),
                          ("pose_mdm", dict(pose_path=motion)), ("train", dict())):
        with pytest.raises(GomError, match="no CPU path"):
            EF.render(eval_type, proc, ck, 64, dev, str(tmp_path / "out"), n_subdivisions=1, batch=2, **kw)
        assert os.path.isdir(os.path.join(str(tmp_path / "out"), "eval", eval_type))
