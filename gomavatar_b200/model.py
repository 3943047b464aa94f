"""``Model``: the per-frame forward of GoMAvatar on the B200-native kernels, behind the reference's signature.

Mirrors reference ``models/model.py::Model`` (constructor ``Model(model_cfg, canonical_info)``, ``forward`` signature at
:184-188, return value ``(rgbs, masks, outputs)`` at :303) and keeps its state-dict keys and SoA parameter layouts
(``vertices [3,V]``, ``so3 [3,F]``, ``scale [3,F]``, ``appearance_module.appearance [3,F]``, buffers ``faces``,
``lbs_weights [J+1,V]``) so reference checkpoints load unchanged.

What differs from the reference, by design:
* every frame of the batch is processed by the same launches (the reference asserts B == 1, gaussian.py:24);
* LBS -> face frame -> covariance -> splat is ~10 kernel launches with no host sync instead of ~150 launches and 6 syncs;
* RGB and alpha are rendered in ONE 4-channel pass instead of the reference's two 3-channel passes
  (gaussian.py:77-94) — same values, same gradients (tests/test_model_gpu.py).
The pose-refinement / non-rigid MLPs are plain torch modules (modules.py); the mesh normal renderer (mesh_renderer.py,
csrc/mesh_raster.cu) and the shadow MLP (shadow.py, csrc/shadow_mlp.cu: tcgen05) are the SURVEY.md §8f "next" rows and
are built from the reference's cfg nodes; all of them can be replaced by callables with the reference's signatures.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from .camera import camera_from_KE
from .rasterizer import rasterize_gaussians
from .skinning import apply_lbs, face_gaussians, get_global_RTs


def _get(cfg, path, default):
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key, None) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


def default_model_cfg(img_size=(512, 512), sigma=1e-3):
    """The subset of the reference's ``cfg.model`` node the hot path reads (configs/default.yaml:56-76)."""
    return SimpleNamespace(img_size=list(img_size), eval_mode=False,
                           canonical_geometry=SimpleNamespace(sigma=sigma, radius_scale=1.0, deform_scale=True, deform_so3=True),
                           appearance=SimpleNamespace(color_init=0.5))


class _Rodrigues(torch.autograd.Function):
    """csrc/rodrigues.cu: the nine entries and their backward in one launch each (torch: ~50 + ~150 elementwise kernels)."""

    @staticmethod
    def forward(ctx, rvec, group, prepend_identity):
        from ._lib import GomRodriguesArgs, call, ptr
        r = rvec.detach().contiguous().float()
        n = r.shape[0]
        per = group + int(prepend_identity)
        R = torch.empty(n // group, per, 3, 3, dtype=torch.float32, device=r.device)
        ctx.args = dict(n_rot=n, group=group, prepend_identity=int(prepend_identity), eps=1e-5)
        call("gom_rodrigues_forward", GomRodriguesArgs(rvec=ptr(r), R=ptr(R), **ctx.args))
        ctx.save_for_backward(r)
        return R

    @staticmethod
    def backward(ctx, g):
        from ._lib import GomRodriguesArgs, call, ptr
        (r,) = ctx.saved_tensors
        gc = g.contiguous().float()
        gr = torch.empty_like(r)
        call("gom_rodrigues_backward", GomRodriguesArgs(rvec=ptr(r), g_R=ptr(gc), g_rvec=ptr(gr), **ctx.args))
        return gr, None, None


def rodrigues_grouped(rvec, group, prepend_identity=False):
    """rvec [N,3] (N a multiple of ``group``) -> [N / group, group (+1), 3, 3]; with ``prepend_identity`` every group starts with
    the identity (the unrefined root joint of reference models/modules/pose_refinement_module.py:39-48)."""
    if rvec.is_cuda:
        return _Rodrigues.apply(rvec, int(group), bool(prepend_identity))
    Rs = rodrigues(rvec).view(-1, group, 3, 3)
    if prepend_identity:
        eye = torch.eye(3, device=Rs.device, dtype=Rs.dtype)[None, None].expand(Rs.shape[0], 1, 3, 3)
        Rs = torch.cat([eye, Rs], dim=1)
    return Rs


def rodrigues(rvec):
    """reference utils/network_util.py:66-92 (RodriguesModule), theta = sqrt(1e-5 + |r|^2); rvec [B,3] -> [B,3,3]."""
    if rvec.is_cuda:
        return _Rodrigues.apply(rvec, 1, False).view(-1, 3, 3)
    theta = torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))
    r = rvec / theta[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = r[:, 0], r[:, 1], r[:, 2]
    oc = 1.0 - c
    return torch.stack((x * x + (1.0 - x * x) * c, x * y * oc - z * s, x * z * oc + y * s,
                        x * y * oc + z * s, y * y + (1.0 - y * y) * c, y * z * oc - x * s,
                        x * z * oc - y * s, y * z * oc + x * s, z * z + (1.0 - z * z) * c), dim=1).view(-1, 3, 3)


def mesh_edges(faces, vertices):
    """(edge lengths [E] in PyTorch3D ``edges_packed`` order — unique (min, max) vertex pairs sorted by min * V + max —
    and the pairs of faces that share an edge [E', 2]) of reference model.py:115-134.  Like the reference's
    ``range(max_edge_id)`` loop, the edge with the largest id is left out of the connectivity."""
    V = vertices.shape[0]
    e = np.concatenate([faces[:, [1, 2]], faces[:, [2, 0]], faces[:, [0, 1]]], axis=0)
    e = np.sort(e, axis=1)
    key = e[:, 0].astype(np.int64) * V + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    ev = np.stack([uniq // V, uniq % V], axis=1)
    length = np.linalg.norm(vertices[ev[:, 0]] - vertices[ev[:, 1]], axis=1).astype(np.float32)
    F = faces.shape[0]
    face_of = np.tile(np.arange(F), 3)
    order = np.lexsort((face_of, inv))                     # by edge id, then face index (torch.nonzero order)
    eid, fid = inv[order], face_of[order]
    first = np.flatnonzero(np.r_[True, eid[1:] != eid[:-1]])
    count = np.diff(np.r_[first, len(eid)])
    keep = (count == 2) & (eid[first] < uniq.shape[0] - 1)
    conn = np.stack([fid[first[keep]], fid[first[keep] + 1]], axis=1).astype(np.int64)
    return length, conn


def _shadow_module(sub):
    """The tcgen05 shadow MLP (shadow.FusedShadowModule) for the configuration family the reference ships; the plain
    torch module (cuBLAS) for anything else (other widths, skip connections inside the depth)."""
    from . import modules, shadow
    try:
        m = shadow.FusedShadowModule(sub)
        if len(m._linears()) - 1 > shadow.MAX_TRAIN_DEPTH:            # deeper MLPs: inference only on the fused path
            raise NotImplementedError
        return m
    except NotImplementedError:
        return modules.ShadowModule(sub)


class AppearanceModule(nn.Module):
    """reference models/modules/appearance_module.py:6-23 — per-face RGB + zero background buffer."""

    def __init__(self, n_faces, color_init=0.5):
        super().__init__()
        self.appearance = nn.Parameter(torch.ones(3, n_faces) * color_init)
        self.register_buffer("bg_col", torch.zeros(3))

    def forward(self, **kwargs):
        return self.appearance, self.bg_col

    def set(self, appearance):
        """reference appearance_module.py:22-23 — used by ``Model.subdivide``."""
        self.appearance = nn.Parameter(appearance)


class Model(nn.Module):
    def __init__(self, model_cfg, canonical_info, pose_refinement_module=None, non_rigid_module=None,
                 normal_renderer=None, shadow_module=None, strict_raster=True):
        super().__init__()
        self.cfg = model_cfg
        faces = torch.as_tensor(np.asarray(canonical_info["faces"]).astype(np.int64))
        self.register_buffer("faces", faces)
        w = torch.as_tensor(np.asarray(canonical_info["canonical_lbs_weights"])).float().transpose(1, 0)     # [J,V]
        self.register_buffer("lbs_weights", torch.cat([w, torch.zeros_like(w[:1])], dim=0).contiguous())   # [J+1,V]
        F = faces.shape[0]
        verts = torch.as_tensor(np.asarray(canonical_info["canonical_vertex"])).float().transpose(1, 0).contiguous()
        self.vertices = nn.Parameter(verts)
        radius = float(_get(model_cfg, "canonical_geometry.radius_scale", 1.0))
        if _get(model_cfg, "canonical_geometry.deform_so3", True):
            self.so3 = nn.Parameter(torch.zeros(3, F))
        else:
            self.register_buffer("so3", torch.zeros(3, F))
        if _get(model_cfg, "canonical_geometry.deform_scale", True):
            self.scale = nn.Parameter(torch.ones(3, F) * radius)
        else:
            self.register_buffer("scale", torch.ones(3, F) * radius)
        self.appearance_module = AppearanceModule(F, float(_get(model_cfg, "appearance.color_init", 0.5)))
        # The modules around the hot path (reference model.py:88-113): explicit arguments win; otherwise they are built
        # from the reference's own cfg nodes (`name != 'none'`) with this package's implementations.
        from . import mesh_renderer, modules
        def build(given, node, ctor):
            if given is not None:
                return given
            sub = _get(model_cfg, node, None)
            return ctor(sub) if sub is not None and _get(sub, "name", "none") != "none" else None
        self.pose_refinement_module = build(pose_refinement_module, "pose_refinement", modules.PoseRefinementModule)
        self.non_rigid_module = build(non_rigid_module, "non_rigid", modules.NonRigidModule)
        self.normal_renderer = build(normal_renderer, "normal_renderer", lambda sub: mesh_renderer.Renderer(
            sub, canonical_info, img_size=_get(model_cfg, "img_size", [512, 512])))
        self.shadow_module = build(shadow_module, "shadow_module", _shadow_module)
        tel, conn = mesh_edges(faces.numpy(), verts.numpy().T)
        self.register_buffer("target_edge_length", torch.from_numpy(tel))       # reference model.py:58-60,127-134
        self.register_buffer("face_connectivity", torch.from_numpy(conn), persistent=False)
        self.strict_raster = strict_raster
        self.last_raster_aux = None             # {'status', 'tile_offset', 'inst_capacity'} of the latest forward
        self.keep_raster_aux = False            # True: keep every intermediate raster buffer there (tests / debugging)
        # The mesh branch (vertex normals -> normal map + soft silhouette -> shadow MLP) and the splat branch (face Gaussians ->
        # rasterizer) both start from the posed vertices and meet at `rgbs = albedos * shadings`: on a CUDA device the mesh branch
        # can run on a side stream (forward and, through autograd, backward); inside a CUDA-graph capture the side stream forks from /
        # joins the capturing stream.  Measured (bench.py --full-model): 2.78 -> 2.75 ms at one frame, 10.53 -> 10.60 ms at 8 frames —
        # within the noise, so it is off by default.
        self.mesh_side_stream = False
        self._side_streams = {}

    def forward(self, K, E, cnl_gtfms, dst_Rs, dst_Ts, dst_posevec=None, canonical_joints=None,
                i_iter=1e7, bgcolor=None, global_R=None, global_T=None, tb=None):
        B = dst_Rs.shape[0]
        W, H = _get(self.cfg, "img_size", [512, 512])
        sigma = float(_get(self.cfg, "canonical_geometry.sigma", 1e-3))

        if self.pose_refinement_module is not None and i_iter >= _get(self.cfg, "pose_refinement.kick_in_iter", 0):
            J = dst_Rs.shape[1]                                    # reference model.py:193-196
            delta = self.pose_refinement_module(dst_posevec)
            dst_Rs = torch.matmul(dst_Rs.reshape(B * J, 3, 3), delta.reshape(B * J, 3, 3)).reshape(B, J, 3, 3)

        vertices_canonical = self.vertices
        if self.non_rigid_module is not None and i_iter >= _get(self.cfg, "non_rigid.kick_in_iter", 0):
            vertices_pose, _, _ = self.non_rigid_module(vertices_canonical.unsqueeze(0), dst_posevec, i_iter, R=None, S=None)
        else:                                                      # reference model.py:199-210
            vertices_pose = vertices_canonical.unsqueeze(0)

        global_Rs, global_Ts = get_global_RTs(cnl_gtfms, dst_Rs, dst_Ts)
        vertices_observation = apply_lbs(vertices_pose, global_Rs, global_Ts, self.lbs_weights)     # [B,3,V]
        if global_R is not None:                                   # reference model.py:218-221 (train_pose.py)
            Rg = rodrigues(global_R.reshape(-1, 3))
            gT = global_T.reshape(-1, 3)
            vertices_observation = Rg @ vertices_observation + gT[:, :, None]

        normal = normal_mask = shadings = None
        mesh_branch = self.normal_renderer is not None and self.shadow_module is not None
        side = None
        if mesh_branch and self.mesh_side_stream and vertices_observation.is_cuda:
            dev = vertices_observation.device
            if dev not in self._side_streams:
                self._side_streams[dev] = torch.cuda.Stream(device=dev)
            side, cur = self._side_streams[dev], torch.cuda.current_stream(dev)
            side.wait_stream(cur)                                  # fork
            with torch.cuda.stream(side):
                normal, normal_mask, shadings = self._mesh_branch(vertices_observation, K, E, B, H, W)
                for t_ in (normal, normal_mask, shadings):
                    if t_ is not None:
                        t_.record_stream(cur)
            vertices_observation.record_stream(side)

        means3D, cov3D = face_gaussians(vertices_observation, self.faces, self.so3, self.scale, sigma)
        appearance, bg_feat = self.appearance_module()
        colors = torch.cat([appearance.permute(1, 0), torch.ones_like(appearance[:1]).permute(1, 0)], dim=1)   # [F,4]
        opacity = torch.ones(B, colors.shape[0], device=colors.device)
        view, proj, tanfov = camera_from_KE(K, E, H, W)
        bg = torch.cat([bg_feat, bg_feat.new_zeros(1)])[None].expand(B, 4)
        aux = {}
        rgba, radii, final_T, _ = rasterize_gaussians(means3D, cov3D, colors, opacity, view, proj, tanfov, bg, H, W,
                                                      interleaved=True, strict=self.strict_raster, aux=aux,
                                                      color_grad_channels=3)
        # what a training loop checks lazily (overflow flag, N_dup); everything else would only pin ~100 MB until the next call
        self.last_raster_aux = aux if self.keep_raster_aux else {k: aux[k] for k in ("status", "tile_offset", "inst_capacity")}
        albedos, masks = rgba[..., :3], rgba[..., 3]

        if mesh_branch:                                                            # reference model.py:271-287
            if side is not None:
                torch.cuda.current_stream(vertices_observation.device).wait_stream(side)      # join
            else:
                normal, normal_mask, shadings = self._mesh_branch(vertices_observation, K, E, B, H, W)
            from .losses import shade_rgba                  # albedo * shading and the alpha channel, one launch each way
            rgbs, masks = shade_rgba(rgba, shadings)
        else:
            rgbs = albedos

        outputs = {}
        if self.training:
            outputs["colors"] = appearance.permute(1, 0)
            outputs["face_connectivity"] = self.face_connectivity
            outputs["target_edge_length"] = self.target_edge_length
            outputs["vertices_observation"] = vertices_observation
            # reference model.py:223-224,295-296: PyTorch3D Meshes for its loss function; here light torch views with the
            # same accessors (meshes.py) — nothing is computed unless a caller asks (regularizers.compute_loss does not)
            from .meshes import Meshes
            outputs["mesh"] = Meshes(vertices_observation.permute(0, 2, 1), self.faces)
            outputs["mesh_canonical"] = Meshes(vertices_canonical.permute(1, 0)[None], self.faces)
            outputs["albedo"] = albedos[0]
            outputs["radii"] = radii
            if normal is not None:
                outputs["normal"], outputs["normal_mask"], outputs["shadow"] = normal, normal_mask[..., 0], shadings
        return rgbs, masks, outputs

    def active_param_groups(self, i_iter):
        """Names of the ``get_param_groups`` groups whose parameters take part in ``forward(..., i_iter=i_iter)`` — what
        ``dist.ArenaAdam.step(active=...)`` needs to reproduce torch.optim.Adam, which skips parameters whose ``.grad`` is
        None: the pose-refinement and non-rigid MLPs before their kick_in_iter (reference models/model.py:193-210)."""
        names = ["appearance", "canonical_geometry_xyz", "canonical_geometry", "shadow"]
        if self.pose_refinement_module is not None and i_iter >= _get(self.cfg, "pose_refinement.kick_in_iter", 0):
            names.append("pose_refinement")
        if self.non_rigid_module is not None and i_iter >= _get(self.cfg, "non_rigid.kick_in_iter", 0):
            names.append("non_rigid")
        return names

    def subdivide(self, need_face_connectivity=True):
        """reference models/model.py:136-179: split every face into 4 at its edge midpoints (subdivision.py mirrors
        utils/pc_util.py::subdivide).  New vertices take the mean position and the mean LBS weights of their edge's end
        points; per-face parameters (so3, scale, colour) are repeated for the 4 children; ``target_edge_length`` and
        ``face_connectivity`` are rebuilt (``need_face_connectivity=False``, eval.py:305, leaves zeros like the
        reference).  Every trainable tensor is a NEW ``nn.Parameter`` afterwards, so the caller re-creates its optimizer
        (train.py:343-346) — and its ``dist.FlatArena`` — exactly as with the reference.  Deterministic host code: all
        ranks of a frame-sharded run call it at the same iteration and stay identical without communication.
        One deviation: the reference's ``target_edge_length`` turns float64 here (``torch.tensor`` of trimesh's float64
        vertices, model.py:174-177); it stays float32 like before the subdivision."""
        from .subdivision import subdivide_mesh
        dev = self.vertices.device
        v_new, f_new, attrs, _ = subdivide_mesh(self.vertices.detach().T.cpu().numpy(), self.faces.cpu().numpy(),
                                                {"weights": self.lbs_weights.detach().T.cpu().numpy()})
        rep4 = lambda t: t.detach()[..., None].repeat(1, 1, 4).reshape(t.shape[0], -1).contiguous()
        appearance, so3, scale = rep4(self.appearance_module.appearance), rep4(self.so3), rep4(self.scale)
        f_new = f_new.astype(np.int64)
        self.vertices = nn.Parameter(torch.from_numpy(v_new).float().T.contiguous().to(dev))
        self.faces = torch.from_numpy(f_new).to(dev)
        self.lbs_weights = torch.from_numpy(attrs["weights"]).float().T.contiguous().to(dev)
        self.appearance_module.set(appearance)
        self.so3 = nn.Parameter(so3) if isinstance(self.so3, nn.Parameter) else so3
        self.scale = nn.Parameter(scale) if isinstance(self.scale, nn.Parameter) else scale
        tel, conn = mesh_edges(f_new, v_new)
        self.target_edge_length = torch.from_numpy(tel).to(dev)
        self.face_connectivity = torch.from_numpy(conn).to(dev) if need_face_connectivity else \
            torch.zeros(tel.shape[0], 2, dtype=torch.int64, device=dev)
        self.last_raster_aux = None

    def print_info(self):
        """reference models/model.py:181-182."""
        import logging
        logging.info(f"the number of effective points is {self.vertices.shape[1]}")

    def get_lbs_weights(self):
        return self.lbs_weights

    def _mesh_branch(self, vertices_observation, K, E, B, H, W):
        """reference model.py:271-280: camera-space vertex normals -> normal map (+ soft silhouette) -> pseudo-shading"""
        from .mesh_renderer import vertex_normals_cam
        normals = vertex_normals_cam(vertices_observation, self.faces, E)            # model.py:271-273 in one launch each way
        normal, normal_mask = self.normal_renderer(vertices_observation, normals, K, E, faces=self.faces)
        shadings = self.shadow_module(normal.reshape(B, H * W, 3)).reshape(B, H, W, 1) * 2
        return normal, normal_mask, shadings

    def _vertex_normals(self, verts_b3v):
        """PyTorch3D ``Meshes.verts_normals_padded`` semantics: area-weighted face normals
        accumulated on vertices, normalised with eps 1e-6.  Only used when a normal renderer is plugged in."""
        from .mesh_renderer import vertex_normals
        return vertex_normals(verts_b3v.permute(0, 2, 1), self.faces)

    def get_param_groups(self, cfg):
        """reference models/model.py:305-324 (lr names from configs/default.yaml:91-99): the same groups in the same order
        — the frozen ``lbs_weights`` buffer first (lr 0, never receives a gradient), then appearance, vertices, scale, so3,
        and the MLPs that exist — so that ``optimizer.state_dict()`` of a ``torch.optim.Adam`` built from it has the
        reference's layout and ``train.py --resume`` loads either side's checkpoints.  ``dist.ArenaAdam`` ignores group
        members that are not trainable."""
        lr = cfg.lr if hasattr(cfg, "lr") else cfg["lr"]
        def g(k, default=None):
            v = getattr(lr, k, None) if not isinstance(lr, dict) else lr.get(k, None)
            if v is None:
                if default is None:
                    raise KeyError(f"cfg.lr.{k} is missing")
                return default
            return v
        groups = [{"name": "lbs_weights", "params": [self.lbs_weights], "lr": g("lbs_weights", 0.0)},
                  {"name": "appearance", "params": list(self.appearance_module.parameters()), "lr": g("appearance")},
                  {"name": "canonical_geometry_xyz", "params": [self.vertices], "lr": g("canonical_geometry_xyz")},
                  {"name": "canonical_geometry", "params": [self.scale], "lr": g("canonical_geometry")},
                  {"name": "canonical_geometry", "params": [self.so3], "lr": g("canonical_geometry")}]
        for name, mod in (("non_rigid", self.non_rigid_module), ("pose_refinement", self.pose_refinement_module),
                          ("shadow", self.shadow_module)):
            if isinstance(mod, nn.Module):
                groups.append({"name": name, "params": list(mod.parameters()), "lr": g(name)})
        return groups
