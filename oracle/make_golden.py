"""Golden-vector generator (test infrastructure).  Runs ONLY in the build container, where /root/reference exists.

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

It executes the reference's OWN code on seeded synthetic inputs and stores inputs + outputs as small fixtures,
so that the oracle restatement (and through it the CUDA path) is pinned to the reference wherever the
reference is runnable on a CPU:

* ``golden_lbs.npz``      reference ``utils/body_util.py`` ``get_global_RTs`` + ``apply_lbs`` (imported unchanged).
* ``golden_steiner.npz``  reference ``models/model.py::get_transformation_from_triangle_steiner`` (module imported
                          unchanged; third-party imports it cannot satisfy offline are stubbed in sys.modules).
* ``golden_model.npz``    reference ``Model.forward`` + ``Renderer.forward`` UNMODIFIED, end to end on CPU:
                          ``.cuda()`` is patched to a no-op, PyTorch3D's ``Meshes``/``so3_exp_map`` and the
                          ``diff_gaussian_rasterization`` package are stubs (the latter backed by the C oracle and
                          recording exactly what the reference hands to the rasterizer), normal renderer / shadow
                          module are replaced by constants (they are "next" rows).
* ``golden_lpips.npz``    reference ``utils/lpips`` LPIPS(net='vgg', pnet_rand=True) under torch.manual_seed(0)
                          (ImageNet trunk weights are not downloadable offline) + its in-tree linear heads.

Nothing here is imported by the product or by GPU tests; the GPU box has no /root/reference.
"""
from __future__ import annotations

import os
import sys
import types
from typing import NamedTuple

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Meshes:
    """Minimal stand-in for pytorch3d.structures.Meshes (semantics: SURVEY.md App. B)."""

    def __init__(self, verts, faces):
        self.v, self.f = verts[0], faces[0]

    def verts_packed(self):
        return self.v

    def faces_packed(self):
        return self.f

    def _edges(self):
        f = self.f
        e = torch.cat([f[:, [1, 2]], f[:, [2, 0]], f[:, [0, 1]]], dim=0)
        e = torch.sort(e, dim=1)[0]
        V = self.v.shape[0]
        key = e[:, 0] * V + e[:, 1]
        uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
        return torch.stack([uniq // V, uniq % V], dim=1), inv.reshape(3, -1).t()

    def edges_packed(self):
        return self._edges()[0]

    def faces_packed_to_edges_packed(self):
        return self._edges()[1]

    def verts_normals_padded(self):
        v, f = self.v, self.f
        n = torch.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]], dim=1)
        vn = torch.zeros_like(v)
        for k in range(3):
            vn = vn.index_add(0, f[:, k], n)
        return torch.nn.functional.normalize(vn, eps=1e-6, dim=1)[None]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


RECORDED = []


class GaussianRasterizer(torch.nn.Module):
    """Records what the reference passes at gaussian.py:83-91 and answers with the C oracle."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        from oracle import raster
        s = self.raster_settings
        assert shs is None and scales is None and rotations is None
        rec = dict(means3D=means3D.detach().numpy().copy(), colors=colors_precomp.detach().numpy().copy(),
                   opacities=opacities.detach().numpy().copy(), cov6=cov3D_precomp.detach().numpy().copy(),
                   view=s.viewmatrix.detach().numpy().copy(), proj=s.projmatrix.detach().numpy().copy(),
                   tanfovx=s.tanfovx, tanfovy=s.tanfovy, bg=s.bg.detach().numpy().copy(),
                   H=s.image_height, W=s.image_width, campos=s.campos.detach().numpy().copy())
        RECORDED.append(rec)
        out = raster.forward(rec["means3D"], rec["cov6"], rec["colors"], rec["opacities"], rec["view"], rec["proj"],
                             s.tanfovx, s.tanfovy, rec["bg"][:3], s.image_height, s.image_width)
        return torch.from_numpy(out["color"]), torch.from_numpy(out["radii"])


def _install_stubs():
    from oracle import geometry as G
    _stub_module("seaborn", color_palette=lambda *a, **k: [(0, 0, 0)])
    tm = _stub_module("trimesh")
    tm.remesh = _stub_module("trimesh.remesh", faces_to_edges=None, grouping=None)
    p3 = _stub_module("pytorch3d")
    p3.ops = _stub_module("pytorch3d.ops")
    p3.ops.knn = _stub_module("pytorch3d.ops.knn", knn_points=None)
    p3.structures = _stub_module("pytorch3d.structures", Meshes=_Meshes)
    p3.transforms = _stub_module("pytorch3d.transforms")
    p3.transforms.so3 = _stub_module("pytorch3d.transforms.so3", so3_exp_map=G.so3_exp_map, so3_log_map=None)
    p3.loss = _stub_module("pytorch3d.loss")
    p3.loss.chamfer = _stub_module("pytorch3d.loss.chamfer", chamfer_distance=None)
    _stub_module("diff_gaussian_rasterization", GaussianRasterizationSettings=GaussianRasterizationSettings,
                 GaussianRasterizer=GaussianRasterizer)


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference is mounted"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    from gomavatar_b200 import synthetic as S
    _install_stubs()
    torch.Tensor.cuda = lambda self, *a, **k: self            # reference Model.__init__ calls .cuda() (model.py:58,60)

    # ---------------------------------------------------------------- LBS (reference code imported unchanged)
    from utils.body_util import apply_lbs, get_global_RTs      # noqa: E402  (reference)
    scene = S.make_humanoid(2000, seed=0)
    frames = S.make_frames(scene, 3, img_size=64, seed=3)
    params = S.make_params(scene, seed=1)
    t = torch.from_numpy
    Rs, Ts = get_global_RTs(t(frames["cnl_gtfms"]), t(frames["dst_Rs"]), t(frames["dst_Ts"]))
    v_obs = torch.stack([apply_lbs(t(params["vertices"])[None], Rs[b:b + 1], Ts[b:b + 1], t(scene.lbs_weights))[0]
                         for b in range(3)])
    np.savez_compressed(os.path.join(OUT, "golden_lbs.npz"),
                        vertices=params["vertices"], lbs_weights=scene.lbs_weights, cnl_gtfms=frames["cnl_gtfms"],
                        dst_Rs=frames["dst_Rs"], dst_Ts=frames["dst_Ts"],
                        global_Rs=Rs.numpy(), global_Ts=Ts.numpy(), vertices_observation=v_obs.numpy())

    # ---------------------------------------------------------------- Steiner frame + full Model.forward
    cwd = os.getcwd()
    os.chdir(REF)                                             # make_cfg opens a relative path (configs/__init__.py:14)
    try:
        from configs import make_cfg
        import models.model as ref_model                      # reference module, unchanged
        cfg = make_cfg("exps/zju-mocap_377.yaml")
    finally:
        os.chdir(cwd)
    tri = t(v_obs.numpy()[0]).permute(1, 0)[t(scene.faces).reshape(-1)].reshape(scene.n_faces, 3, 3)
    A = ref_model.get_transformation_from_triangle_steiner(tri, 1e-3)
    np.savez_compressed(os.path.join(OUT, "golden_steiner.npz"), triangles=tri.numpy(), sigma=np.float32(1e-3),
                        transform=A.numpy())

    H = W = 64
    cfg.model.img_size = [W, H]
    cfg.model.normal_renderer.name = "none"                   # mesh normal renderer / shadow MLP: "next" rows
    cfg.model.shadow_module.name = "none"
    if "eval_mode" not in cfg.model:
        cfg.model.eval_mode = False
    model = ref_model.Model(cfg.model, scene.canonical_info())
    model.train()
    with torch.no_grad():
        model.so3.copy_(t(params["so3"]))
        model.scale.copy_(t(params["scale"]))
        model.appearance_module.appearance.copy_(t(params["appearance"]))
    model.normal_renderer = lambda verts, normals, K, E, faces=None: (torch.zeros(1, H, W, 3), torch.ones(1, H, W, 1))
    model.shadow_module = lambda n: torch.full((1, H * W, 1), 0.5)
    gold = {}
    for b in range(3):
        RECORDED.clear()
        sl = lambda k: t(frames[k][b:b + 1])
        rgbs, masks, outputs = model(sl("K"), sl("E"), sl("cnl_gtfms"), sl("dst_Rs"), sl("dst_Ts"),
                                     dst_posevec=sl("dst_posevec"), i_iter=0, bgcolor=sl("bgcolor"))
        gold[f"rgbs_{b}"] = rgbs.detach().numpy()
        gold[f"masks_{b}"] = masks.detach().numpy()
        gold[f"albedo_{b}"] = outputs["albedo"].detach().numpy()
        assert len(RECORDED) == 2                             # two 3-channel passes (gaussian.py:82-92)
        for p, rec in enumerate(RECORDED):
            for k in ("means3D", "colors", "opacities", "cov6", "view", "proj", "bg", "campos"):
                gold[f"pass{p}_{k}_{b}"] = rec[k]
            gold[f"pass{p}_tanfov_{b}"] = np.array([rec["tanfovx"], rec["tanfovy"]], np.float64)
    # test-time pose optimisation branch (model.py:218-221)
    RECORDED.clear()
    gR, gT = torch.tensor([0.1, -0.2, 0.05]), torch.tensor([0.02, 0.01, -0.03])
    sl = lambda k: t(frames[k][0:1])
    rgbs, masks, _ = model(sl("K"), sl("E"), sl("cnl_gtfms"), sl("dst_Rs"), sl("dst_Ts"), dst_posevec=sl("dst_posevec"),
                           i_iter=0, global_R=gR, global_T=gT)
    gold["global_R"], gold["global_T"] = gR.numpy(), gT.numpy()
    gold["rigid_means3D"] = RECORDED[0]["means3D"]
    gold["rigid_cov6"] = RECORDED[0]["cov6"]
    gold["rigid_rgbs"] = rgbs.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "golden_model.npz"), n_faces=np.int64(2000), scene_seed=np.int64(0),
                        frames_seed=np.int64(3), params_seed=np.int64(1), img_size=np.int64(64),
                        faces=scene.faces.astype(np.int32), so3=params["so3"], scale=params["scale"],
                        appearance=params["appearance"],
                        **{k: frames[k] for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor")},
                        **gold)

    # ---------------------------------------------------------------- LPIPS (reference utils/lpips, random trunk)
    try:
        from utils import lpips as ref_lpips
        torch.manual_seed(0)
        net = ref_lpips.LPIPS(net="vgg", pnet_rand=True, verbose=False)
        trunk_sum = float(sum(p.double().abs().sum() for p in net.net.parameters()))
        g = torch.Generator().manual_seed(5)
        x0 = torch.rand(2, 3, 64, 64, generator=g)
        x1 = (x0 + 0.1 * torch.randn(2, 3, 64, 64, generator=g)).clamp(0, 1)
        x0r = x0.clone().requires_grad_(True)
        val = net(2 * x0r - 1, 2 * x1 - 1)
        val.sum().backward()
        heads = {f"lin{k}": net.lins[k].model[-1].weight.detach().numpy().reshape(-1) for k in range(5)}
        np.savez_compressed(os.path.join(OUT, "golden_lpips.npz"), x0=x0.numpy(), x1=x1.numpy(),
                            value=val.detach().numpy(), grad_x0=x0r.grad.numpy(), trunk_abs_sum=np.float64(trunk_sum),
                            **heads)
        print("lpips golden:", val.detach().reshape(-1).numpy(), "trunk |w| sum", trunk_sum)
    except Exception as e:  # pragma: no cover
        print("LPIPS golden skipped:", repr(e))
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def modules_golden():
    """``python -m oracle.make_golden modules``: golden_modules.npz — the reference's OWN ShadowModule, NonRigidModule and
    PoseRefinementModule (models/modules/*.py imported unchanged; their pytorch3d imports are stubbed, none is used on
    these paths) built from exps/zju-mocap_377.yaml, every parameter re-drawn from a seeded generator (the reference's
    1e-5 last-layer init would make the outputs trivially small), run on seeded inputs."""
    assert os.path.isdir(REF)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    _install_stubs()
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from configs import make_cfg
        cfg = make_cfg("exps/zju-mocap_377.yaml")
        from models.modules.shadow_module import ShadowModule
        from models.modules.non_rigid_module import NonRigidModule
        from models.modules.pose_refinement_module import PoseRefinementModule
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(11)
    out = {}

    def redraw(mod, prefix):
        for k, v in mod.state_dict().items():
            new = torch.randn(v.shape, generator=g) * (0.3 if v.dim() > 1 else 0.1) / (v.shape[-1] ** 0.5 if v.dim() > 1 else 1.0) * 3
            v.copy_(new)
            out[f"{prefix}.{k}"] = new.numpy()

    with torch.no_grad():
        sh = ShadowModule(cfg.model.shadow_module)
        redraw(sh, "shadow")
        n = torch.nn.functional.normalize(torch.randn(2, 500, 3, generator=g), dim=-1) * torch.rand(2, 500, 1, generator=g) * 3
        n[0, :50] = 0                                          # background pixels carry a zero normal
        out["shadow_in"], out["shadow_out"] = n.numpy(), sh(n).numpy()
        nr = NonRigidModule(cfg.model.non_rigid)
        redraw(nr, "non_rigid")
        xyz = torch.randn(2, 3, 300, generator=g) * 0.5
        pv = torch.randn(2, 69, generator=g) * 0.3
        out["non_rigid_xyz"], out["non_rigid_posevec"] = xyz.numpy(), pv.numpy()
        for it in (150000, 163000, 187500, 300000):
            out[f"non_rigid_out_{it}"] = nr(xyz, pv, it)[0].numpy()
        pr = PoseRefinementModule(cfg.model.pose_refinement)
        redraw(pr, "pose_refinement")
        out["pose_in"] = pv.numpy()
        out["pose_out"] = pr(pv).numpy()
    for name in ("shadow_module", "non_rigid", "pose_refinement"):
        for k, v in dict(cfg.model[name]).items():
            if isinstance(v, (int, float)):
                out[f"cfg.{name}.{k}"] = np.float64(v)
    np.savez_compressed(os.path.join(OUT, "golden_modules.npz"), **out)
    print("golden_modules.npz:", len(out), "arrays")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "modules":
    modules_golden()


def dataset_golden():
    """``golden_dataset.npz``: the reference's OWN reader (``dataset/train.py::Dataset``, imported unchanged) run on the folder
    that ``gomavatar_b200.dataset_io.write_synthetic_dataset`` wrote (oracle/dataset_fixture.py).  OpenCV is absent offline:
    ``cv2`` is a stub whose Rodrigues is exact axis-angle, whose undistort is the identity (the fixture has zero
    distortion) and whose resize refuses to resample (the fixture is read at its native size) — none of them changes a
    value.  ``termcolor`` (imported by utils/image_util.py for log colouring) is an empty stub."""
    import tempfile
    from oracle import dataset_fixture as DF

    def rodrigues(rvec):
        v = np.asarray(rvec, dtype=np.float64).reshape(3)
        th = float(np.linalg.norm(v))
        if th < 1e-12:
            return (np.eye(3), None)
        k = v / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return (np.cos(th) * np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * np.outer(k, k), None)

    def undistort(img, K, D):
        assert not np.any(np.asarray(D) != 0)
        return img

    def resize(img, size, **kw):
        assert size is not None and tuple(size) == (img.shape[1], img.shape[0]), "the fixture is read at its native size"
        return img

    _stub_module("cv2", Rodrigues=rodrigues, undistort=undistort, resize=resize, INTER_LANCZOS4=4, INTER_LINEAR=1)
    _stub_module("termcolor", colored=lambda s, *a, **k: s)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ref_train = importlib.import_module("dataset.train")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        DF.build(tmp)
        ds = ref_train.Dataset(tmp, bgcolor=[255.0, 128.0, 0.0], target_size=[DF.W, DF.H])
        assert len(ds) == DF.N_FRAMES
        for i in range(len(ds)):
            item = ds[i]
            for k, v in item.items():
                if k == "frame_name":
                    out[f"item{i}.frame_name"] = np.array(v)
                else:
                    out[f"item{i}.{k}"] = np.asarray(v)
        info = ds.get_canonical_info()
        for k, v in info.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    out[f"info.{k}.{kk}"] = np.asarray(vv)
            elif v is not None:
                out[f"info.{k}"] = np.asarray(v)
        # the random-crop branch, under a fixed numpy seed
        np.random.seed(3)
        dc = ref_train.Dataset(tmp, bgcolor=[0.0, 0.0, 0.0], target_size=[DF.W, DF.H], crop_size=[32, 24])
        item = dc[1]
        for k in ("K", "target_rgbs", "target_masks"):
            out[f"crop.{k}"] = np.asarray(item[k])
    np.savez_compressed(os.path.join(OUT, "golden_dataset.npz"), **out)
    print("golden_dataset.npz:", len(out), "arrays")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dataset":
    dataset_golden()


def _extract_functions(path, names):
    """Source text of the top-level functions `names` of a reference file, compiled in a namespace the caller fills —
    the reference's own code is EXECUTED from where it lies (nothing is copied into this repository) without importing
    the file's unsatisfiable module-level dependencies (tensorboard, pytorch3d, seaborn, skimage ...)."""
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    assert set(out) == set(names), (set(names) - set(out))
    return out


class _LossMeshes(_Meshes):
    """+ what reference utils/network_util.py::mesh_laplacian_smoothing asks of pytorch3d Meshes (uniform Laplacian
    L = D^-1 A - I as a sparse matrix: pytorch3d/structures/meshes.py::laplacian_packed, restated)."""
    device = torch.device("cpu")

    def isempty(self):
        return False

    def __len__(self):
        return 1

    def num_verts_per_mesh(self):
        return torch.tensor([self.v.shape[0]])

    def verts_packed_to_mesh_idx(self):
        return torch.zeros(self.v.shape[0], dtype=torch.long)

    def laplacian_packed(self):
        e = self.edges_packed()
        V = self.v.shape[0]
        idx = torch.cat([e.t(), e.flip(1).t()], dim=1)                        # both directions
        A = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (V, V)).coalesce()
        deg = torch.sparse.sum(A, dim=1).to_dense()
        w = 1.0 / deg[idx[0]]
        L = torch.sparse_coo_tensor(idx, w, (V, V))
        L = L - torch.sparse_coo_tensor(torch.arange(V).repeat(2, 1), torch.ones(V), (V, V))
        return L.coalesce()


def _p3d_mesh_normal_consistency(meshes):
    """pytorch3d.loss.mesh_normal_consistency (0.7.0, absent offline) restated step by step from its published source,
    independently of gomavatar_b200/regularizers.py: every edge gathers the faces that use it (``faces_packed_to_edges_packed``
    sorted by edge), ``mesh_normal_consistency_find_verts`` lists all pairs (i < j) of those faces per edge, the normal of a
    face relative to the edge (v0, v1) is (v1 - v0) x (its opposite vertex - v0) — obtained as the sum over the face's three
    corners, two of which give zero — and the loss is the mean of 1 - cos(n_i, -n_j) over the pairs."""
    verts, faces = meshes.verts_packed(), meshes.faces_packed()
    edges, face_to_edge = meshes.edges_packed(), meshes.faces_packed_to_edges_packed()
    E, Fn = edges.shape[0], faces.shape[0]
    with torch.no_grad():
        edge_idx = face_to_edge.reshape(Fn * 3)
        vert_idx = faces.view(1, Fn, 3).expand(3, Fn, 3).transpose(0, 1).reshape(3 * Fn, 3)
        edge_idx, sort_idx = edge_idx.sort()
        vert_idx = vert_idx[sort_idx]
        edge_num = edge_idx.bincount(minlength=E)
        pairs, start = [], 0
        for n in edge_num.tolist():                              # mesh_normal_consistency_find_verts (C++)
            for i in range(n):
                for j in range(i + 1, n):
                    pairs.append((start + i, start + j))
            start += n
        pair_idx = torch.tensor(pairs, dtype=torch.int64)
    v0, v1 = verts[edges[edge_idx, 0]], verts[edges[edge_idx, 1]]
    n = sum(torch.cross(v1 - v0, verts[vert_idx[:, k]] - v0, dim=1) for k in range(3))
    n0, n1 = n[pair_idx[:, 0]], -n[pair_idx[:, 1]]
    loss = 1 - torch.cosine_similarity(n0, n1, dim=1)
    return loss.sum() / loss.numel()                             # one mesh: weights = 1 / number of pairs


def loss_golden():
    """``golden_loss.npz``: the reference's OWN ``unpack`` + ``compute_loss`` (train.py:53-55, :98-163), its own
    ``mesh_laplacian_smoothing`` / ``mesh_color_consistency`` (utils/network_util.py:669-799) and its own LPIPS (random trunk
    under seed 0, in-tree heads) on seeded inputs, with the loss coefficients of exps/zju-mocap_377.yaml.  Stand-ins: the
    Meshes object (stub above) and ``pytorch3d.loss.mesh_normal_consistency`` (absent offline; restated above from
    PyTorch3D's published source, independently of the product — that ONE term stays unpinned against PyTorch3D itself)."""
    import types as _t
    import torch.nn.functional as F
    from gomavatar_b200 import regularizers as RG
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.model import mesh_edges
    if REF not in sys.path:
        sys.path.insert(0, REF)
    src_train = _extract_functions(os.path.join(REF, "train.py"), ["unpack", "compute_loss"])
    src_net = _extract_functions(os.path.join(REF, "utils", "network_util.py"), ["mesh_laplacian_smoothing", "mesh_color_consistency"])
    ns = {"torch": torch, "F": F}
    for s_ in src_net.values():
        exec(s_, ns)
    ns["mesh_normal_consistency"] = _p3d_mesh_normal_consistency
    for s_ in src_train.values():
        exec(s_, ns)
    from utils import lpips as ref_lpips
    torch.manual_seed(0)
    lp = ref_lpips.LPIPS(net="vgg", pnet_rand=True, verbose=False)

    H = W = 64
    scene = S.make_humanoid(2000, seed=0)
    g = torch.Generator().manual_seed(21)
    verts = torch.from_numpy(scene.vertices) + 0.004 * torch.randn(scene.n_vertices, 3, generator=g)
    faces = torch.from_numpy(scene.faces).long()
    _, conn = mesh_edges(scene.faces.astype(np.int64), scene.vertices)
    conn = torch.from_numpy(conn)
    mesh = _LossMeshes(verts[None], faces[None])
    mesh.conn = conn
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    blob = (((xx - 30) ** 2 / 1.5 + (yy - 34) ** 2) < 18 ** 2).float()
    rgb_raw = torch.rand(1, H, W, 3, generator=g) * blob[None, ..., None]
    mask_pred = (blob[None] * (0.6 + 0.4 * torch.rand(1, H, W, generator=g))).clamp(0, 1)
    bgcolor = torch.rand(1, 3, generator=g)
    rgb_gt = torch.rand(1, H, W, 3, generator=g)
    mask_gt = (((xx - 32) ** 2 + (yy - 32) ** 2) < 17 ** 2).float()[None]
    normal_mask = (blob[None] * torch.rand(1, H, W, generator=g)).clamp(0, 1)
    colors = torch.rand(scene.n_faces, 3, generator=g)
    cfgl = _t.SimpleNamespace(rgb=_t.SimpleNamespace(coeff=1.0), mask=_t.SimpleNamespace(coeff=5.0), lpips=_t.SimpleNamespace(coeff=1.0),
                              laplacian=_t.SimpleNamespace(coeff_canonical=0.0, coeff_observation=10.0),
                              normal=_t.SimpleNamespace(mask_dilate=True, kernel_size=7, coeff_mask=1.0, coeff_consist=0.10),
                              color_consist=_t.SimpleNamespace(coeff=0.050))
    outputs = {"mesh": mesh, "mesh_canonical": mesh, "normal_mask": normal_mask, "colors": colors, "face_connectivity": conn}
    with torch.no_grad():
        rgb = ns["unpack"](rgb_raw, mask_pred, bgcolor)
        total, losses = ns["compute_loss"](rgb, mask_pred, outputs, rgb_gt, mask_gt, cfgl, None, 0, lpips_func=lp)
    out = {"verts": verts.numpy(), "faces": scene.faces.astype(np.int64), "face_connectivity": conn.numpy(), "rgb_raw": rgb_raw.numpy(),
           "mask_pred": mask_pred.numpy(), "bgcolor": bgcolor.numpy(), "rgb_gt": rgb_gt.numpy(), "mask_gt": mask_gt.numpy(),
           "normal_mask": normal_mask.numpy(), "colors": colors.numpy(), "total": np.float64(total)}
    for k, v in losses.items():
        out[f"unscaled.{k}"] = np.float64(v["unscaled"])
        out[f"scaled.{k}"] = np.float64(v["scaled"])
    np.savez_compressed(os.path.join(OUT, "golden_loss.npz"), **out)
    print("golden_loss.npz:", {k: float(v["unscaled"]) for k, v in losses.items()}, "total", float(total))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "loss":
    loss_golden()


# ------------------------------------------------------------------------------------------------ mesh subdivision
def _tm_faces_to_edges(faces, return_index=False):
    """trimesh.geometry.faces_to_edges (third party, absent offline; restated from its published source)."""
    faces = np.asanyarray(faces)
    return faces[:, [0, 1, 1, 2, 2, 0]].reshape((-1, 2))


class _tm_grouping:
    """trimesh.grouping.hashable_rows / unique_rows for small integer rows (restated; trimesh is unpinned in
    requirements.txt:9 and this row hash is the same in 3.x and 4.x)."""

    @staticmethod
    def hashable_rows(data, digits=None):
        as_int = np.asanyarray(data).astype(np.int64)
        precision = int(np.floor(64 / as_int.shape[1]))
        assert np.abs(as_int).max() < 2 ** (precision - 1)
        hashable = np.zeros(len(as_int), dtype=np.int64)
        for offset, column in enumerate(as_int.T):
            np.bitwise_xor(hashable, column << (offset * precision), out=hashable)
        return hashable

    @staticmethod
    def unique_rows(data, digits=None):
        rows = _tm_grouping.hashable_rows(data, digits=digits)
        _, unique, inverse = np.unique(rows, return_index=True, return_inverse=True)
        return unique, inverse


class _Trimesh:
    """What utils/pc_util.py::subdivide asks of trimesh.Trimesh (process=True merges coincident vertices: none here)."""

    def __init__(self, vertices, faces, vertex_attributes=None):
        self.vertices = np.asanyarray(vertices, dtype=np.float64)
        self.faces = np.asanyarray(faces, dtype=np.int64)
        self.vertex_attributes = dict(vertex_attributes or {})

    @property
    def edges(self):
        return _tm_faces_to_edges(self.faces)


def subdivide_golden():
    """``golden_subdivide.npz``: the reference's OWN ``Model.subdivide`` (models/model.py:136-179, module imported
    unchanged) calling its OWN ``utils/pc_util.py::subdivide`` / ``_subdivide``, on a seeded 2000-face humanoid with random
    per-face parameters: the complete state dict after one subdivision, after a second one with
    ``need_face_connectivity=False`` (eval.py:305; stored as SHA-256 digests of the arrays, the comparison is bit-exact
    anyway), and the face connectivity.  Stand-ins: PyTorch3D ``Meshes`` (stub
    above) and the three trimesh pieces restated above."""
    assert os.path.isdir(REF)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from gomavatar_b200 import synthetic as S
    _install_stubs()
    tm = sys.modules["trimesh"]
    tm.Trimesh = _Trimesh
    tm.remesh.faces_to_edges, tm.remesh.grouping = _tm_faces_to_edges, _tm_grouping
    torch.Tensor.cuda = lambda self, *a, **k: self
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from configs import make_cfg
        import models.model as ref_model
        cfg = make_cfg("exps/zju-mocap_377.yaml")
    finally:
        os.chdir(cwd)
    cfg.model.img_size = [64, 64]
    cfg.model.normal_renderer.name = "none"
    cfg.model.shadow_module.name = "none"
    cfg.model.non_rigid.name = "none"
    cfg.model.pose_refinement.name = "none"
    if "eval_mode" not in cfg.model:
        cfg.model.eval_mode = False
    scene = S.make_humanoid(2000, seed=4)
    model = ref_model.Model(cfg.model, scene.canonical_info())
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        model.so3.copy_(0.1 * torch.randn(model.so3.shape, generator=g))
        model.scale.copy_(0.7 + 0.6 * torch.rand(model.scale.shape, generator=g))
        model.appearance_module.appearance.copy_(torch.rand(model.so3.shape, generator=g))
        model.vertices.add_(1e-3 * torch.randn(model.vertices.shape, generator=g))
    out = {"seed_scene": np.int64(4), "n_faces": np.int64(2000),
           "in.faces": model.faces.numpy().copy(), "in.face_connectivity": model.face_connectivity.numpy().copy()}
    for k, v in model.state_dict().items():
        out[f"in.{k}"] = v.numpy().copy()
    model.subdivide()
    for k, v in model.state_dict().items():
        out[f"s1.{k}"] = v.detach().numpy().copy()
    out["s1.face_connectivity"] = model.face_connectivity.numpy().copy()
    model.subdivide(need_face_connectivity=False)
    import hashlib
    for k in ("vertices", "faces", "lbs_weights"):
        a = np.ascontiguousarray(model.state_dict()[k].detach().numpy())
        out[f"s2.{k}.shape"] = np.asarray(a.shape, np.int64)
        out[f"s2.{k}.dtype"] = np.asarray(str(a.dtype))
        out[f"s2.{k}.sha256"] = np.asarray(hashlib.sha256(a.tobytes()).hexdigest())
    out["s2.face_connectivity_shape"] = np.asarray(model.face_connectivity.shape, np.int64)
    np.savez_compressed(os.path.join(OUT, "golden_subdivide.npz"), **out)
    print("golden_subdivide.npz:", {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith("s1.")})


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "subdivide":
    subdivide_golden()


def dataset_cv2_golden():
    """``golden_dataset_cv2.npz``: the reference's OWN reader with the REAL OpenCV (importable in this container: cv2
    4.13) on the fixture folder, for what the stubbed run above cannot pin — ``cv2.undistort`` with non-zero lens
    distortion (frames 1, 2), ``cv2.resize`` INTER_LANCZOS4 / INTER_LINEAR to a ``target_size`` different from the files',
    and the ``resize_img_scale`` (0.5, 0.5) branch taken without ``target_size``.  Only ``termcolor`` is a stub."""
    import pickle
    import tempfile
    import cv2  # noqa: F401  (the real one)
    from oracle import dataset_fixture as DF
    _stub_module("termcolor", colored=lambda s, *a, **k: s)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ref_train = importlib.import_module("dataset.train")
    out = {"cv2_version": np.array(cv2.__version__)}
    with tempfile.TemporaryDirectory() as tmp:
        DF.build(tmp)
        DF.add_distortion(tmp)
        for tag, kw in (("resized", dict(target_size=[64, 56])), ("halved", dict())):
            ds = ref_train.Dataset(tmp, bgcolor=[255.0, 128.0, 0.0], **kw)
            for i in range(len(ds)):
                item = ds[i]
                for k in ("K", "E", "target_rgbs", "target_masks", "bgcolor"):
                    out[f"{tag}.item{i}.{k}"] = np.asarray(item[k])
    np.savez_compressed(os.path.join(OUT, "golden_dataset_cv2.npz"), **out)
    print("golden_dataset_cv2.npz:", len(out), "arrays;", {k: out[k].shape for k in ("resized.item1.target_rgbs", "halved.item1.target_rgbs")})


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dataset_cv2":
    dataset_cv2_golden()


def dataset_zju_views_golden():
    """``golden_dataset_zju_views.npz``: the reference's OWN novel-view / novel-pose reader (``dataset/test.py::Dataset``,
    imported unchanged, real OpenCV) on the synthetic raw ZJU-MoCap capture of oracle/dataset_fixture.py::build_raw_zju."""
    import tempfile
    import cv2
    from oracle import dataset_fixture as DF
    _stub_module("termcolor", colored=lambda s, *a, **k: s)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ref_test = importlib.import_module("dataset.test")
    out = {"cv2_version": np.array(cv2.__version__)}
    with tempfile.TemporaryDirectory() as tmp:
        raw, proc = os.path.join(tmp, "raw"), os.path.join(tmp, "processed")
        os.makedirs(raw)
        DF.build_raw_zju(raw, proc)
        for tag, kw in (("view", dict(test_type="view", skip=3, exclude_view=0)),
                        ("pose", dict(test_type="pose", skip=1, exclude_training_view=False))):
            ds = ref_test.Dataset(raw, proc, bgcolor=[10.0, 200.0, 90.0], **kw)
            out[f"{tag}.len"] = np.int64(len(ds))
            for i in range(len(ds)):
                item = ds[i]
                for k, v in item.items():
                    out[f"{tag}.item{i}.{k}"] = np.array(v) if k == "frame_name" else np.asarray(v)
        info = ds.get_canonical_info()
        out["info.canonical_vertex"] = np.asarray(info["canonical_vertex"])
    np.savez_compressed(os.path.join(OUT, "golden_dataset_zju_views.npz"), **out)
    print("golden_dataset_zju_views.npz:", len(out), "arrays; view items", int(out["view.len"]), "pose items", int(out["pose.len"]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dataset_zju_views":
    dataset_zju_views_golden()


def dataset_freeview_golden():
    """``golden_dataset_freeview.npz``: the reference's OWN free-view reader (``dataset/freeview.py::Dataset``, imported
    unchanged, real OpenCV) on the fixture folder: 7 cameras on a circle for frame 1, both rotation conventions."""
    import tempfile
    import cv2  # noqa: F401
    from oracle import dataset_fixture as DF
    _stub_module("termcolor", colored=lambda s, *a, **k: s)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ref_fv = importlib.import_module("dataset.freeview")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        DF.build(tmp)
        for tag, kw in (("zju", dict(src_type="zju_mocap", target_size=[DF.W, DF.H])), ("wild", dict(src_type="wild", bgcolor=[0.0, 64.0, 255.0]))):
            ds = ref_fv.Dataset(tmp, 1, total_frames=7, **kw)
            out[f"{tag}.len"] = np.int64(len(ds))
            for i in range(len(ds)):
                for k, v in ds[i].items():
                    out[f"{tag}.item{i}.{k}"] = np.array(v) if k == "frame_name" else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "golden_dataset_freeview.npz"), **out)
    print("golden_dataset_freeview.npz:", len(out), "arrays")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dataset_freeview":
    dataset_freeview_golden()


def dataset_newpose_golden():
    """``golden_dataset_newpose.npz``: the reference's OWN motion-file reader (``dataset/newpose.py::Dataset``, imported
    unchanged; eval.py --type pose_mdm) on the fixture folder with a seeded 3-pose MDM-format file (the reference opens
    images/frame_<idx>.png for every pose, so the motion cannot be longer than the fixture)."""
    import tempfile
    from oracle import dataset_fixture as DF
    _stub_module("termcolor", colored=lambda s, *a, **k: s)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ref_np = importlib.import_module("dataset.newpose")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        DF.build(tmp)
        pose_file = DF.write_mdm_motion(os.path.join(tmp, "motion.npy"), DF.N_FRAMES)
        ds = ref_np.Dataset(tmp, pose_file)
        out["len"] = np.int64(len(ds))
        for i in range(len(ds)):
            for k, v in ds[i].items():
                if k in ("target_rgbs", "target_masks"):
                    out[f"item{i}.{k}.shape"] = np.asarray(np.asarray(v).shape)
                    assert not np.any(v)
                else:
                    out[f"item{i}.{k}"] = np.array(v) if k == "frame_name" else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "golden_dataset_newpose.npz"), **out)
    print("golden_dataset_newpose.npz:", len(out), "arrays")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dataset_newpose":
    dataset_newpose_golden()
