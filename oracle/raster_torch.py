"""Oracle cross-check (test infrastructure): an independent, differentiable torch restatement of the splat
rasterizer for SMALL cases, in any dtype (float64 for gradient checks).

Purpose: prove that ``raster_oracle.c``'s hand-written backward (App. A.6/A.7) is the derivative of its
forward.  Vectorised over pixels, sequential over depth-sorted Gaussians.  Two upstream conventions are kept
on purpose: the 0.99 alpha clamp is transparent to the gradient, and integer decisions (cull, radius, tile
rect, skip tests, termination) carry no gradient.
"""
from __future__ import annotations

import math

import torch


def render(means3D, cov6, colors, opacity, view, proj, tanfovx, tanfovy, bg, H, W):
    dt = means3D.dtype
    P, C = colors.shape
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    view = view.to(dt).reshape(4, 4)
    proj = proj.to(dt).reshape(4, 4)
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], dim=1)
    pv = hom @ view            # row-vector convention (App. A.2)
    ph = hom @ proj
    pw = 1.0 / (ph[:, 3] + 1e-7)
    ppx, ppy = ph[:, 0] * pw, ph[:, 1] * pw
    tz = pv[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tx = torch.clamp(pv[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(pv[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], -1),
                     torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], -1)], dim=1)      # [P,2,3]
    Rm = view[:3, :3].T                                                                       # E[:3,:3]
    M = J @ Rm[None]
    S = torch.stack([torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2]], -1),
                     torch.stack([cov6[:, 1], cov6[:, 3], cov6[:, 4]], -1),
                     torch.stack([cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1)], dim=1)
    c2 = M @ S @ M.transpose(1, 2)
    a, b, c = c2[:, 0, 0] + 0.3, c2[:, 0, 1], c2[:, 1, 1] + 0.3
    det = a * c - b * b
    conA, conB, conC = c / det, -b / det, a / det
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((ppx + 1.0) * W - 1.0) * 0.5
    py = ((ppy + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pxd, pyd = px.detach(), py.detach()
    minx = torch.clamp(torch.trunc((pxd - radius) / 16), 0, gx)
    miny = torch.clamp(torch.trunc((pyd - radius) / 16), 0, gy)
    maxx = torch.clamp(torch.trunc((pxd + radius + 15) / 16), 0, gx)
    maxy = torch.clamp(torch.trunc((pyd + radius + 15) / 16), 0, gy)
    visible = (tz.detach() > 0.2) & (det.detach() != 0) & ((maxx - minx) * (maxy - miny) > 0)

    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    tile_x, tile_y = torch.floor(xs / 16), torch.floor(ys / 16)
    T = torch.ones(H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    acc = torch.zeros(C, H, W, dtype=dt)
    order = sorted(range(P), key=lambda i: (float(tz[i]), i))
    for i in order:
        if not bool(visible[i]):
            continue
        in_rect = (tile_x >= minx[i]) & (tile_x < maxx[i]) & (tile_y >= miny[i]) & (tile_y < maxy[i])
        dx, dy = px[i] - xs, py[i] - ys
        power = -0.5 * (conA[i] * dx * dx + conC[i] * dy * dy) - conB[i] * dx * dy
        raw = opacity[i] * torch.exp(power)
        alpha = raw + (torch.clamp(raw, max=0.99) - raw).detach()    # clamp is gradient-transparent
        ok = in_rect & (~done) & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = ok & (test_T.detach() < 1e-4)
        done = done | stop
        ok = ok & (~stop)
        w = torch.where(ok, alpha * T, torch.zeros_like(T))
        acc = acc + colors[i][:, None, None] * w[None]
        T = torch.where(ok, test_T, T)
    return acc + T[None] * bg.to(dt)[:C, None, None], T
