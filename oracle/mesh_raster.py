"""Oracle (test infrastructure): PyTorch3D 0.7.0 mesh rasterisation as the reference's normal-map renderer uses it
(reference models/modules/renderer/mesh.py:23-128, utils/pc_util.py:11-46, models/model.py:271-274), restated in
plain torch on the CPU, naive O(pixels x faces), differentiable by autograd.

PARITY UNPINNED: pytorch3d is not installed here and is not part of /root/reference; the semantics below are restated
from its published sources (rasterize_meshes.cu `CheckPixelInsideFace`, geometry_utils.cuh, blending.py
`sigmoid_alpha_blend` / `hard_rgb_blend`, structures/meshes.py `verts_normals_packed`) and SURVEY.md App. B.

* ``ndc_T_world``      utils/pc_util.py:30-46 (NDC with +X left, +Y up; the shorter image side spans [-1, 1])
* ``vertex_normals``   Meshes.verts_normals_packed: area-weighted face normals accumulated on vertices, normalised (eps 1e-6)
* ``rasterize``        per pixel, every face: bounding-box test with sqrt(blur_radius) margin, zero-area cull (1e-8),
                       barycentric coordinates (area + 1e-8), pz = bary . z, cull pz < 0, squared distance to the nearest
                       edge segment, signed negative inside, keep if inside or dist < blur_radius; the K faces of
                       smallest pz survive.
* ``normal_map``       NormalShader: sum of the hit face's three vertex normals (bary weights are ones), background 0
* ``soft_silhouette``  SoftSilhouetteShader: 1 - prod_k (1 - sigmoid(-dist_k / 1e-4)) over the K = 50 nearest faces
"""
from __future__ import annotations

import math

import torch

K_EPS = 1e-8


def ndc_T_world(xyz_world, K, E, H, W):
    """xyz_world [V,3], K [3,3], E [4,4] -> [V,3] (x_ndc, y_ndc, z_cam)."""
    cam = xyz_world @ E[:3, :3].T + E[:3, 3]
    uv = cam @ K.T
    xy = uv[:, :2] / uv[:, 2:]
    if H < W:
        xs = -((xy[:, 0] / H) * 2.0 - (W / H))
        ys = -((xy[:, 1] / H) * 2.0 - 1.0)
    else:
        xs = -((xy[:, 0] / W) * 2.0 - 1.0)
        ys = -((xy[:, 1] / W) * 2.0 - (H / W))
    return torch.stack([xs, ys, cam[:, 2]], dim=-1)


def vertex_normals(verts, faces):
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    n = torch.zeros_like(verts)
    n = n.index_add(0, faces[:, 1], torch.cross(v2 - v1, v0 - v1, dim=1))
    n = n.index_add(0, faces[:, 2], torch.cross(v0 - v2, v1 - v2, dim=1))
    n = n.index_add(0, faces[:, 0], torch.cross(v1 - v0, v2 - v0, dim=1))
    return torch.nn.functional.normalize(n, eps=1e-6, dim=1)


def pixel_centers_ndc(H, W, dtype=torch.float32):
    """PyTorch3D's pixel grid: pixel (yi, xi) samples NDC (sx - (2 xi + 1)/S, sy - (2 yi + 1)/S), S = min(H, W)."""
    S = min(H, W)
    xs = W / S - (2.0 * torch.arange(W, dtype=dtype) + 1.0) / S
    ys = H / S - (2.0 * torch.arange(H, dtype=dtype) + 1.0) / S
    return xs, ys


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist2(px, py, ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    t = ((px - ax) * dx + (py - ay) * dy) / torch.where(l2 > K_EPS, l2, torch.ones_like(l2))
    tt = t.clamp(0.0, 1.0)
    qx, qy = ax + tt * dx, ay + tt * dy
    d_seg = (px - qx) ** 2 + (py - qy) ** 2
    d_pt = (px - bx) ** 2 + (py - by) ** 2
    return torch.where(l2 > K_EPS, d_seg, d_pt)


def rasterize(verts_ndc, faces, H, W, blur_radius=0.0, faces_per_pixel=1, chunk_rows=8):
    """-> pix_to_face [H,W,K] int64 (-1 = none), zbuf [H,W,K], dists [H,W,K] (differentiable wrt verts_ndc xy)."""
    dt = verts_ndc.dtype
    K = faces_per_pixel
    xs, ys = pixel_centers_ndc(H, W, dt)
    v = verts_ndc[faces]                                   # [F,3,3]
    ax, ay, az = v[:, 0, 0], v[:, 0, 1], v[:, 0, 2]
    bx, by, bz = v[:, 1, 0], v[:, 1, 1], v[:, 1, 2]
    cx, cy, cz = v[:, 2, 0], v[:, 2, 1], v[:, 2, 2]
    br = math.sqrt(blur_radius)
    xmin, xmax = torch.minimum(torch.minimum(ax, bx), cx) - br, torch.maximum(torch.maximum(ax, bx), cx) + br
    ymin, ymax = torch.minimum(torch.minimum(ay, by), cy) - br, torch.maximum(torch.maximum(ay, by), cy) + br
    z_invalid = torch.maximum(torch.maximum(az, bz), cz) < K_EPS
    face_area = _edge(cx, cy, ax, ay, bx, by)              # EdgeFunction(v2, v0, v1) == EdgeFunction(v0, v1, v2) up to order
    zero_area = face_area.abs() <= K_EPS
    p2f = torch.full((H, W, K), -1, dtype=torch.int64)
    zb = torch.full((H, W, K), -1.0, dtype=dt)
    ds = torch.full((H, W, K), -1.0, dtype=dt)
    INF = torch.tensor(float("inf"), dtype=dt)
    for r0 in range(0, H, chunk_rows):
        r1 = min(H, r0 + chunk_rows)
        py = ys[r0:r1][:, None, None]
        px = xs[None, :, None]
        outside_bb = (px > xmax) | (px < xmin) | (py > ymax) | (py < ymin) | z_invalid
        area = face_area + K_EPS
        w0 = _edge(px, py, bx, by, cx, cy) / area
        w1 = _edge(px, py, cx, cy, ax, ay) / area
        w2 = _edge(px, py, ax, ay, bx, by) / area
        pz = w0 * az + w1 * bz + w2 * cz
        dist = torch.minimum(torch.minimum(_seg_dist2(px, py, ax, ay, bx, by), _seg_dist2(px, py, ax, ay, cx, cy)),
                             _seg_dist2(px, py, bx, by, cx, cy))
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        keep = ~outside_bb & ~zero_area & (pz >= 0) & (inside | (dist < blur_radius))
        signed = torch.where(inside, -dist, dist)
        zkey = torch.where(keep, pz.detach(), INF)
        k = min(K, zkey.shape[-1])
        zsel, idx = torch.topk(zkey, k, dim=-1, largest=False, sorted=True)
        valid = torch.isfinite(zsel)
        p2f[r0:r1, :, :k] = torch.where(valid, idx, torch.full_like(idx, -1))
        zb[r0:r1, :, :k] = torch.where(valid, torch.gather(pz, -1, idx), torch.full_like(zsel, -1.0))
        ds[r0:r1, :, :k] = torch.where(valid, torch.gather(signed, -1, idx), torch.full_like(zsel, -1.0))
    return p2f, zb, ds


def normal_map(pix_to_face, faces, vert_normals):
    """NormalShader + hard_rgb_blend + the reference's multiplication by alpha: [H,W,3], 0 on the background."""
    f = pix_to_face[..., 0]
    hit = f >= 0
    fn = vert_normals[faces].sum(dim=1)                    # weights are ones, not barycentric (mesh.py:23-30)
    out = torch.zeros(f.shape + (3,), dtype=vert_normals.dtype)
    out[hit] = fn[f[hit]]
    return out


def soft_silhouette(pix_to_face, dists, sigma=1e-4):
    mask = (pix_to_face >= 0).to(dists.dtype)
    prob = torch.sigmoid(-dists / sigma) * mask
    return 1.0 - torch.prod(1.0 - prob, dim=-1)


def render(xyz_world, faces, K, E, H, W, training=True, sigma_cfg=1e-5, faces_per_pixel=50):
    """reference mesh.py::Renderer.forward + models/model.py:271-274 for one frame: (normal [H,W,3], mask [H,W] | None)."""
    ndc = ndc_T_world(xyz_world, K, E, H, W)
    vn = vertex_normals(xyz_world, faces) @ E[:3, :3].T
    p2f, _, _ = rasterize(ndc, faces, H, W, 0.0, 1)
    nm = normal_map(p2f, faces, vn)
    if not training:
        return nm, None
    blur = math.log(1.0 / 1e-4 - 1.0) * sigma_cfg
    p2f_s, _, d_s = rasterize(ndc, faces, H, W, blur, faces_per_pixel)
    return nm, soft_silhouette(p2f_s, d_s)
