"""Normal-map renderer of the posed mesh on libgom_b200.so — drop-in for reference
``models/modules/renderer/mesh.py::Renderer`` (constructor ``Renderer(module_cfg, canonical_info)``, ``forward(
xyzs_observation, vertex_normals, K, E, faces)`` -> ``(normal_map * alpha [B,H,W,3], soft mask [B,H,W,1] | None)``),
which the reference builds on PyTorch3D's naive mesh rasterizer.

``ndc_T_world`` (reference utils/pc_util.py:11-46) stays a handful of differentiable torch ops; the rasterisation, the
shading, the soft silhouette and their backward are csrc/mesh_raster.cu.  ``vertex_normals`` restates PyTorch3D's
``Meshes.verts_normals_padded`` (what models/model.py:271 calls) for callers that do not have PyTorch3D.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import GomMeshRasterArgs, GomNdcArgs, GomVertexNormalsArgs, call, ptr


def ndc_T_world(xyzs_world, K, E, H, W):
    """reference utils/pc_util.py:30-46.  xyzs_world [B,3,V], K [B,3,3], E [B,4,4] -> [B,V,3] (x_ndc, y_ndc, z_cam)."""
    ones = torch.ones_like(xyzs_world[:, :1])
    cam_ = torch.bmm(E, torch.cat([xyzs_world, ones], dim=1))
    cam = cam_[:, :3] / cam_[:, 3:]
    xys_ = torch.bmm(K, cam)
    xys = xys_[:, :2] / xys_[:, 2:]
    if H < W:
        xs = -((xys[:, 0, :] / H) * 2. - (W / H))
        ys = -((xys[:, 1, :] / H) * 2. - 1.)
    else:
        xs = -((xys[:, 0, :] / W) * 2. - 1.)
        ys = -((xys[:, 1, :] / W) * 2. - (H / W))
    return torch.stack([xs, ys, cam[:, 2]], dim=-1)


def vertex_normals(verts_bv3, faces):
    """PyTorch3D ``Meshes.verts_normals_padded``: area-weighted face normals accumulated on the vertices, eps 1e-6."""
    f = faces.long()
    v0, v1, v2 = (verts_bv3.index_select(1, f[:, k]) for k in range(3))      # backward = index_add, no sort
    n = torch.zeros_like(verts_bv3)
    n = n.index_add(1, f[:, 1], torch.cross(v2 - v1, v0 - v1, dim=-1))
    n = n.index_add(1, f[:, 2], torch.cross(v0 - v2, v1 - v2, dim=-1))
    n = n.index_add(1, f[:, 0], torch.cross(v1 - v0, v2 - v0, dim=-1))
    return torch.nn.functional.normalize(n, eps=1e-6, dim=-1)


class _NdcTWorld(torch.autograd.Function):
    """``ndc_T_world`` as one launch each way (csrc/mesh_prep.cu); K and E are inputs of the step and get no gradient."""

    @staticmethod
    def forward(ctx, xyzs_world, K, E, H, W):
        B, _, V = xyzs_world.shape
        v, Kc, Ec = xyzs_world.detach().contiguous().float(), K.detach().contiguous().float(), E.detach().contiguous().float()
        out = torch.empty(B, V, 3, dtype=torch.float32, device=v.device)
        call("gom_ndc_forward", GomNdcArgs(n_frames=B, n_verts=V, height=H, width=W, verts=ptr(v), K=ptr(Kc), E=ptr(Ec), ndc=ptr(out)))
        ctx.save_for_backward(v, Kc, Ec)
        ctx.hw = (H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        v, Kc, Ec = ctx.saved_tensors
        B, _, V = v.shape
        gc = g.contiguous().float()
        gv = torch.empty_like(v)
        call("gom_ndc_backward", GomNdcArgs(n_frames=B, n_verts=V, height=ctx.hw[0], width=ctx.hw[1], verts=ptr(v), K=ptr(Kc),
                                            E=ptr(Ec), dL_dndc=ptr(gc), dL_dverts=ptr(gv)))
        return gv, None, None, None, None


class _VertexNormalsCam(torch.autograd.Function):
    """PyTorch3D ``verts_normals_padded`` of the posed mesh rotated into the camera frame (reference models/model.py:271-273):
    two launches each way (csrc/mesh_prep.cu) instead of ~15 + ~25 gather / cross / index_add / normalize kernels."""

    @staticmethod
    def forward(ctx, verts_b3v, faces, E):
        B, _, V = verts_b3v.shape
        v, Ec = verts_b3v.detach().contiguous().float(), E.detach().contiguous().float()
        fc = faces.contiguous()
        if fc.dtype not in (torch.int32, torch.int64):
            fc = fc.long()
        acc = torch.empty(B, V, 3, dtype=torch.float32, device=v.device)
        out = torch.empty(B, V, 3, dtype=torch.float32, device=v.device)
        call("gom_vertex_normals_forward", GomVertexNormalsArgs(n_frames=B, n_verts=V, n_faces=fc.shape[0], faces_int64=int(fc.dtype == torch.int64),
                                                                verts=ptr(v), faces=ptr(fc), E=ptr(Ec), acc=ptr(acc), normals_cam=ptr(out)))
        ctx.save_for_backward(v, fc, Ec, acc)
        return out

    @staticmethod
    def backward(ctx, g):
        v, fc, Ec, acc = ctx.saved_tensors
        B, _, V = v.shape
        gc = g.contiguous().float()
        scratch, gv = torch.empty_like(acc), torch.empty_like(v)
        call("gom_vertex_normals_backward", GomVertexNormalsArgs(n_frames=B, n_verts=V, n_faces=fc.shape[0], faces_int64=int(fc.dtype == torch.int64),
                                                                 verts=ptr(v), faces=ptr(fc), E=ptr(Ec), acc=ptr(acc), dL_dnormals_cam=ptr(gc),
                                                                 scratch=ptr(scratch), dL_dverts=ptr(gv)))
        return gv, None, None


def vertex_normals_cam(verts_b3v, faces, E):
    """[B,3,V] posed vertices, faces [F,3], E [B,4,4] -> camera-space unit vertex normals [B,V,3] (CUDA kernels; the torch
    formulation ``vertex_normals`` + bmm below is the readable definition and the CPU path of the tests)."""
    if verts_b3v.is_cuda:
        return _VertexNormalsCam.apply(verts_b3v, faces, E)
    n = vertex_normals(verts_b3v.permute(0, 2, 1), faces)
    return torch.bmm(E[:, :3, :3], n.permute(0, 2, 1)).permute(0, 2, 1)


MESH_BIN = 8                 # pixels per side of a binning tile (csrc/mesh_raster.cu kBin)


def default_list_capacity(n_faces):
    return max(8 * int(n_faces), 1 << 16)


class _MeshRaster(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts_ndc, vert_normals, faces, H, W, soft, blur_radius, faces_per_pixel, capacity, aux, strict=True):
        if verts_ndc.device.type != "cuda":
            raise _lib.GomError("mesh renderer: inputs must live on a CUDA device (no CPU path exists)")
        B, V, _ = verts_ndc.shape
        F = faces.shape[0]
        dev = verts_ndc.device
        vn, nn_ = verts_ndc.detach().contiguous().float(), vert_normals.detach().contiguous().float()
        fc = faces.contiguous()
        if fc.dtype not in (torch.int32, torch.int64):
            fc = fc.long()
        T = ((W + MESH_BIN - 1) // MESH_BIN) * ((H + MESH_BIN - 1) // MESH_BIN)
        cap = int(capacity) if capacity else default_list_capacity(F)
        e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        while True:
            st = dict(tile_count=e(B, T, dtype=torch.int32), tile_offset=e(B, T + 1, dtype=torch.int32),
                      tile_cursor=e(B, T, dtype=torch.int32), face_list=e(B, cap, dtype=torch.int32),
                      status=e(B, dtype=torch.int32), worklist=e(B * T + 4, dtype=torch.int32),
                      pix_to_face=e(B, H, W, dtype=torch.int32), normal_map=e(B, H, W, 3))
            if soft:
                st.update(alpha=e(B, H, W), zcut=e(B, H, W), idcut=e(B, H, W, dtype=torch.int32))
            a = GomMeshRasterArgs(n_frames=B, n_verts=V, n_faces=F, height=H, width=W, faces_int64=int(fc.dtype == torch.int64),
                                  soft=int(bool(soft)), faces_per_pixel=int(faces_per_pixel), blur_radius=float(blur_radius),
                                  list_capacity=cap, verts_ndc=ptr(vn), faces=ptr(fc), vert_normals=ptr(nn_),
                                  **{k: ptr(v) for k, v in st.items()})
            call("gom_mesh_raster_forward", a)
            if not strict or torch.cuda.is_current_stream_capturing():      # no host sync: overflow stays a device flag in aux["status"]
                break
            if int(st["status"].max().item()) & _lib.STATUS_OVERFLOW:       # binning lists too small: regrow, like the splat path
                need = int(st["tile_offset"][:, T].to(torch.int64).bitwise_and(0xFFFFFFFF).max().item())
                cap = int(need * 1.25) + 1024
                continue
            break
        ctx.meta = (B, V, F, H, W, bool(soft), float(blur_radius), int(faces_per_pixel), cap)
        saved = [vn, nn_, fc, st["tile_count"], st["tile_offset"], st["tile_cursor"], st["face_list"], st["status"],
                 st["pix_to_face"], st["normal_map"], st["worklist"]]
        if soft:
            saved += [st["alpha"], st["zcut"], st["idcut"]]
        ctx.save_for_backward(*saved)
        if aux is not None:
            aux.update(st)
        ctx.mark_non_differentiable(st["pix_to_face"])
        alpha = st["alpha"] if soft else torch.zeros(0, device=dev)
        return st["normal_map"], alpha, st["pix_to_face"]

    @staticmethod
    def backward(ctx, g_normal, g_alpha, _g_p2f):
        B, V, F, H, W, soft, blur, K, cap = ctx.meta
        t = ctx.saved_tensors
        vn, nn_, fc, tile_count, tile_offset, tile_cursor, face_list, status, p2f, nmap, worklist = t[:11]
        alpha, zcut, idcut = (t[11], t[12], t[13]) if soft else (None, None, None)
        dev = vn.device
        gn = None if g_normal is None else g_normal.contiguous().float()
        ga = None if (g_alpha is None or not soft) else g_alpha.contiguous().float()
        d_verts = torch.empty(B, V, 3, dtype=torch.float32, device=dev)
        d_vn = torch.empty(B, V, 3, dtype=torch.float32, device=dev)
        a = GomMeshRasterArgs(n_frames=B, n_verts=V, n_faces=F, height=H, width=W, faces_int64=int(fc.dtype == torch.int64),
                              soft=int(soft), faces_per_pixel=K, blur_radius=blur, list_capacity=cap, verts_ndc=ptr(vn),
                              faces=ptr(fc), vert_normals=ptr(nn_), tile_count=ptr(tile_count), tile_offset=ptr(tile_offset),
                              tile_cursor=ptr(tile_cursor), face_list=ptr(face_list), status=ptr(status), worklist=ptr(worklist), pix_to_face=ptr(p2f),
                              normal_map=ptr(nmap), alpha=ptr(alpha), zcut=ptr(zcut), idcut=ptr(idcut), dL_dnormal_map=ptr(gn),
                              dL_dalpha=ptr(ga), dL_dverts_ndc=ptr(d_verts), dL_dvert_normals=ptr(d_vn))
        call("gom_mesh_raster_backward", a)
        return d_verts, d_vn, None, None, None, None, None, None, None, None, None


def rasterize_mesh(verts_ndc, vert_normals, faces, image_height, image_width, soft=False, blur_radius=0.0,
                   faces_per_pixel=50, capacity=None, aux=None, strict=True):
    """verts_ndc [B,V,3], vert_normals [B,V,3], faces [F,3] -> (normal_map [B,H,W,3] with 0 on the background,
    alpha [B,H,W] (empty unless soft), pix_to_face [B,H,W] int32)."""
    return _MeshRaster.apply(verts_ndc, vert_normals, faces, int(image_height), int(image_width), bool(soft),
                             float(blur_radius), int(faces_per_pixel), capacity, aux, bool(strict))


class Renderer(nn.Module):
    """reference models/modules/renderer/mesh.py:64-128.  ``module_cfg`` needs ``img_size`` (W, H); optional ``eval_mode``,
    ``sigma`` (default 1e-4 there; the reference configs set 1e-5, exps/zju-mocap_377.yaml:89)."""

    def __init__(self, module_cfg=None, canonical_info=None, img_size=None, sigma=None, faces_per_pixel=50, **kwargs):
        super().__init__()
        get = lambda k, d: (module_cfg.get(k, d) if isinstance(module_cfg, dict) else getattr(module_cfg, k, d)) \
            if module_cfg is not None else d
        self.img_size = list(img_size if img_size is not None else get("img_size", [512, 512]))
        self.sigma = float(sigma if sigma is not None else get("sigma", 1e-4))
        self.blur_radius = math.log(1. / 1e-4 - 1.) * self.sigma
        self.faces_per_pixel = int(faces_per_pixel)
        self.strict, self.capacity = True, None      # strict=False: never read the overflow flag back (CUDA-graph capture)
        self.last_aux = None

    def forward(self, xyzs_observation, vertex_normals, K, E, faces, **kwargs):
        W, H = self.img_size
        xyzs_ndc = _NdcTWorld.apply(xyzs_observation, K, E, int(H), int(W)) if xyzs_observation.is_cuda else ndc_T_world(xyzs_observation, K, E, H, W)
        B = xyzs_ndc.shape[0]
        vn = vertex_normals if vertex_normals.dim() == 3 else vertex_normals[None]
        if vn.shape[0] != B:
            vn = vn.expand(B, -1, -1)
        aux = {}
        normal, alpha, _ = rasterize_mesh(xyzs_ndc, vn, faces, H, W, soft=self.training, blur_radius=self.blur_radius,
                                          faces_per_pixel=self.faces_per_pixel, aux=aux, capacity=self.capacity, strict=self.strict)
        self.last_aux = aux
        if not self.training:
            return normal, None
        return normal, alpha[..., None]
