"""Mesh subdivision (reference models/model.py:136-179, utils/pc_util.py:49-172) — host logic, runs without a GPU.

Pinned by ``tests/golden/golden_subdivide.npz``: the reference's OWN ``Model.subdivide`` executed from its files in the
build container (``oracle/make_golden.py::subdivide_golden``; stand-ins there: PyTorch3D ``Meshes`` and three small
trimesh helpers).  Integer / index results and every copied or averaged float are compared bit-exactly."""
import hashlib
import os

import numpy as np
import pytest
import torch

from gomavatar_b200 import synthetic as S
from gomavatar_b200.model import Model, default_model_cfg
from gomavatar_b200.subdivision import faces_to_edges, subdivide_mesh, subdivided_sizes, unique_rows

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_subdivide.npz")


def _model_from_golden(g):
    info = {"faces": g["in.faces"], "canonical_vertex": g["in.vertices"].T.copy(),
            "canonical_lbs_weights": g["in.lbs_weights"][:-1].T.copy()}
    m = Model(default_model_cfg((64, 64)), info)
    with torch.no_grad():
        m.so3.copy_(torch.from_numpy(g["in.so3"]))
        m.scale.copy_(torch.from_numpy(g["in.scale"]))
        m.appearance_module.appearance.copy_(torch.from_numpy(g["in.appearance_module.appearance"]))
    return m


def test_model_subdivide_equals_the_reference_model_subdivide():
    g = np.load(GOLDEN)
    m = _model_from_golden(g)
    assert np.array_equal(m.face_connectivity.numpy(), g["in.face_connectivity"])
    # (in.vertices were jittered after the reference constructor ran, so the INITIAL edge lengths are not comparable)
    m.subdivide()
    sd = m.state_dict()
    assert sorted(sd) == sorted(k[3:] for k in g.files if k.startswith("s1.") and k != "s1.face_connectivity")
    for k in ("vertices", "faces", "lbs_weights", "so3", "scale", "appearance_module.appearance", "appearance_module.bg_col"):
        assert sd[k].dtype == torch.from_numpy(g[f"s1.{k}"]).dtype, k
        assert np.array_equal(sd[k].numpy(), g[f"s1.{k}"]), k                      # bit-exact, incl. the vertex numbering
    assert np.array_equal(m.face_connectivity.numpy(), g["s1.face_connectivity"])
    # the reference's buffer turns float64 here (documented deviation: ours stays float32)
    assert sd["target_edge_length"].dtype == torch.float32
    np.testing.assert_allclose(sd["target_edge_length"].numpy(), g["s1.target_edge_length"], rtol=2e-7)
    assert isinstance(m.vertices, torch.nn.Parameter) and isinstance(m.so3, torch.nn.Parameter)
    assert isinstance(m.scale, torch.nn.Parameter) and isinstance(m.appearance_module.appearance, torch.nn.Parameter)

    m.subdivide(need_face_connectivity=False)                                      # eval.py:302-305
    sd = m.state_dict()
    for k in ("vertices", "faces", "lbs_weights"):
        a = np.ascontiguousarray(sd[k].numpy())
        assert list(a.shape) == list(g[f"s2.{k}.shape"]) and str(a.dtype) == str(g[f"s2.{k}.dtype"]), k
        assert hashlib.sha256(a.tobytes()).hexdigest() == str(g[f"s2.{k}.sha256"]), k
    assert list(m.face_connectivity.shape) == list(g["s2.face_connectivity_shape"])
    assert int(m.face_connectivity.abs().sum()) == 0


def test_trimesh_helper_restatements():
    f = np.array([[0, 1, 2], [2, 1, 3]])
    assert faces_to_edges(f).tolist() == [[0, 1], [1, 2], [2, 0], [2, 1], [1, 3], [3, 2]]
    e = np.sort(faces_to_edges(f), axis=1)
    u, inv = unique_rows(e)
    # distinct rows ordered by (max vertex, min vertex); `u` = first occurrence of each
    assert e[u].tolist() == [[0, 1], [0, 2], [1, 2], [1, 3], [2, 3]]
    assert np.array_equal(e[u][inv], e)
    assert u.tolist() == [0, 2, 1, 4, 5]


@pytest.mark.parametrize("n_faces", [2000, 13776])
def test_subdivision_properties(n_faces):
    sc = S.make_humanoid(n_faces, seed=1)
    v, f = sc.vertices.astype(np.float64), sc.faces.astype(np.int64)
    w = np.ascontiguousarray(sc.lbs_weights.T.astype(np.float32))                 # [V,25] like Model.subdivide passes it
    v2, f2, attrs, edges, index = subdivide_mesh(v, f, {"weights": w}, return_edges=True)
    E = 3 * len(f) // 2
    assert v2.shape == (len(v) + E, 3) and f2.shape == (4 * len(f), 3) and edges.shape == (12 * len(f), 2)
    assert subdivided_sizes(len(v), len(f)) == (len(v2), len(f2))
    assert np.array_equal(v2[:len(v)], v) and np.array_equal(attrs["weights"][:len(v)], w)
    assert sorted(index) == list(range(len(f))) and index[len(f) - 1].tolist() == list(range(4 * len(f) - 4, 4 * len(f)))
    # still a closed manifold: every undirected edge is used by exactly two faces, in opposite directions
    d = faces_to_edges(f2)
    key = np.sort(d, axis=1)
    _, counts = np.unique(key[:, 0] * len(v2) + key[:, 1], return_counts=True)
    assert (counts == 2).all() and len(counts) == 3 * len(f2) // 2
    assert len(np.unique(d[:, 0] * len(v2) + d[:, 1])) == len(d)
    # geometry: the 4 children tile their parent with the same orientation (areas add up, normals agree)
    def area_normals(vv, ff):
        n = np.cross(vv[ff[:, 1]] - vv[ff[:, 0]], vv[ff[:, 2]] - vv[ff[:, 0]])
        return n
    n_old, n_new = area_normals(v, f), area_normals(v2, f2).reshape(len(f), 4, 3)
    np.testing.assert_allclose(n_new.sum(1), n_old, rtol=0, atol=1e-12)
    np.testing.assert_allclose(n_new, np.repeat(n_old[:, None] / 4, 4, axis=1), rtol=0, atol=1e-12)
    # every midpoint is the mean of an old edge; its skinning weights are the mean of the end points' (sum stays 1)
    np.testing.assert_allclose(attrs["weights"].sum(1), 1.0, atol=2e-6)
    mid_of = {}
    for a, b, m in np.concatenate([np.stack([f[:, i], f[:, (i + 1) % 3], f2.reshape(len(f), 4, 3)[:, 3, i]], 1) for i in range(3)]):
        mid_of.setdefault(m, (a, b))
    m_idx = np.array(sorted(mid_of))
    ab = np.array([mid_of[m] for m in m_idx])
    assert np.array_equal(m_idx, np.arange(len(v), len(v2)))
    assert np.array_equal(v2[m_idx], (v[ab[:, 0]] + v[ab[:, 1]]) / 2)
    assert np.array_equal(attrs["weights"][m_idx], ((w[ab[:, 0]] + w[ab[:, 1]]) / np.float32(2)))


def test_reference_config_sizes():
    """ZJU 13 776 -> 55 104 faces (27 554 vertices); two levels give 110 210 / 220 416."""
    assert subdivided_sizes(6890, 13776) == (27554, 55104)
    assert subdivided_sizes(6890, 13776, levels=2) == (110210, 220416)


def test_special_attributes_and_refusals():
    v = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    f = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    so3 = np.ones((4, 3), np.float32)
    scale = np.arange(12, dtype=np.float32).reshape(4, 3)
    v2, f2, attrs, _ = subdivide_mesh(v, f, {"so3": so3, "scale": scale})
    assert v2.shape == (10, 3) and f2.shape == (16, 3)
    assert (attrs["so3"][4:] == 0).all() and np.array_equal(attrs["so3"][:4], so3)      # pc_util.py:138-139
    e = np.sort(faces_to_edges(f), axis=1)
    u, _ = unique_rows(e)
    np.testing.assert_allclose(attrs["scale"][4:, 0], np.linalg.norm(scale[e[u][:, 1]] - scale[e[u][:, 0]], axis=1))
    with pytest.raises(ValueError, match="duplicate vertex"):
        subdivide_mesh(np.vstack([v, v[:1]]), f)
    with pytest.raises(ValueError, match="out of range"):
        subdivide_mesh(v, f + 1)
    with pytest.raises(ValueError, match="rows for"):
        subdivide_mesh(v, f, {"weights": np.zeros((3, 2))})
    with pytest.raises(ValueError):
        subdivide_mesh(v[:, :2], f)


def test_resume_replays_the_subdivision_and_loads_strictly(tmp_path):
    """train.py:275-279 — build from the canonical mesh, replay ``subdivide()``, ``load_state_dict`` (strict)."""
    from gomavatar_b200.dataset_io import save_checkpoint
    from gomavatar_b200.dist import FlatArena
    sc = S.make_humanoid(2000, seed=2)
    a = Model(default_model_cfg((64, 64)), sc.canonical_info())
    a.subdivide()
    with torch.no_grad():
        a.vertices.add_(0.01)
        a.so3.normal_()
    path = os.path.join(tmp_path, "iter_7.pt")
    save_checkpoint(path, a, optimizer_state=None, n_iter=7)
    ckpt = torch.load(path, weights_only=False)
    b = Model(default_model_cfg((64, 64)), sc.canonical_info())
    with pytest.raises(RuntimeError):
        b.load_state_dict(ckpt["network"])                                         # shapes differ before the replay
    b.subdivide()
    b.load_state_dict(ckpt["network"])
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    assert torch.equal(a.face_connectivity, b.face_connectivity)
    # every trainable tensor is a new Parameter: optimizers / arenas are rebuilt by the caller (train.py:343-346)
    lr = {"lr": {"appearance": 1e-3, "canonical_geometry_xyz": 1e-4, "canonical_geometry": 1e-3}}
    groups = b.get_param_groups(lr)
    assert [g["name"] for g in groups] == ["lbs_weights", "appearance", "canonical_geometry_xyz", "canonical_geometry", "canonical_geometry"]
    n = sum(p.numel() for grp in groups for p in grp["params"] if p.requires_grad)
    assert n == 3 * b.vertices.shape[1] + 9 * b.faces.shape[0]
    arena = FlatArena(b)
    assert arena.numel == n and b.vertices.data_ptr() == arena.data.data_ptr() + 4 * arena.slices[
        [id(p) for p in arena.params].index(id(b.vertices))][0]


def test_param_groups_have_the_reference_optimizer_layout():
    """models/model.py:305-324: 8 groups with the full ZJU config, the frozen lbs_weights buffer first — so a
    ``torch.optim.Adam`` state dict written by either side loads on the other (train.py:281)."""
    from types import SimpleNamespace as NS
    sc = S.make_humanoid(2000, seed=2)
    cfg = default_model_cfg((64, 64))
    cfg.pose_refinement = NS(name="mlp", embedding_size=69, mlp_width=256, mlp_depth=4, kick_in_iter=0)
    try:
        m = Model(cfg, sc.canonical_info())
    except Exception:                                              # module cfg fields differ: the 5 core groups are enough here
        m = Model(default_model_cfg((64, 64)), sc.canonical_info())
    lr = NS(lr=NS(lbs_weights=0.0, appearance=5e-3, canonical_geometry_xyz=5e-5, canonical_geometry=5e-4, non_rigid=5e-5,
                  pose_refinement=5e-5, shadow=5e-4))
    groups = m.get_param_groups(lr)
    assert [g["name"] for g in groups][:5] == ["lbs_weights", "appearance", "canonical_geometry_xyz", "canonical_geometry", "canonical_geometry"]
    assert groups[0]["params"][0] is m.lbs_weights and groups[3]["params"][0] is m.scale and groups[4]["params"][0] is m.so3
    opt = torch.optim.Adam(groups, betas=(0.9, 0.999))
    w0 = m.lbs_weights.clone()
    (m.vertices.sum() + m.scale.sum()).backward()
    opt.step()
    assert torch.equal(m.lbs_weights, w0)                          # no gradient, lr 0: untouched
    sd = opt.state_dict()
    assert [g["params"] for g in sd["param_groups"]][:5] == [[0], [1], [2], [3], [4]]
    opt2 = torch.optim.Adam(m.get_param_groups(lr), betas=(0.9, 0.999))
    opt2.load_state_dict(sd)
    with pytest.raises(KeyError):
        m.get_param_groups({"lr": {"appearance": 1e-3}})
