"""Development micro-benchmark: rasterizer forward / backward per-launch times with CUDA events (not bench.py)."""
import sys, os, json
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import numpy as np
import torch
from util_scene import raster_inputs
from gomavatar_b200.rasterizer import rasterize_gaussians


def run(n_faces, img, B, iters=20):
    d1 = raster_inputs(n_faces=n_faces, img=img, n_frames=min(B, 4), channels=4)
    dev = torch.device("cuda:0")
    rep = lambda a: np.concatenate([a] * ((B + a.shape[0] - 1) // a.shape[0]), 0)[:B]
    g = lambda k: torch.from_numpy(np.ascontiguousarray(rep(d1[k]))).to(dev)
    m, cv, op, view, proj, tf, bg = (g(k) for k in ("means3D", "cov6", "opacity", "view", "proj", "tanfov", "bg"))
    col = torch.from_numpy(d1["colors"]).to(dev)
    m.requires_grad_(True); cv.requires_grad_(True); col.requires_grad_(True)
    H, W = d1["H"], d1["W"]
    dL = torch.randn(B, H, W, 4, device=dev)
    aux = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf_, tb_ = [], []
    for it in range(iters + 3):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        color, radii, T, nc = rasterize_gaussians(m, cv, col, op, view, proj, tf, bg, H, W, interleaved=True, strict=False, aux=aux)
        e1.record()
        color.backward(dL)
        e2.record()
        torch.cuda.synchronize()
        if it >= 3:
            tf_.append(e0.elapsed_time(e1)); tb_.append(e1.elapsed_time(e2))
        m.grad = cv.grad = col.grad = None
    T_ = ((W + 15) // 16) * ((H + 15) // 16)
    ndup = aux["tile_offset"][:, T_].cpu().numpy().view(np.uint32)
    cnt = aux["tile_count"].cpu().numpy().view(np.uint32)
    print(json.dumps(dict(n_faces=n_faces, img=img, B=B, fwd_ms=float(np.median(tf_)), bwd_ms=float(np.median(tb_)),
                          fwd_us_per_frame=1e3 * float(np.median(tf_)) / B, bwd_us_per_frame=1e3 * float(np.median(tb_)) / B,
                          n_dup_mean=float(ndup.mean()), tiles_nonempty=float((cnt > 0).sum(1).mean()),
                          tile_max=int(cnt.max()), status=int(aux["status"].max()))))


if __name__ == "__main__":
    for nf, img, B in ((30000, 512, 1), (30000, 512, 8), (30000, 512, 32), (13776, 512, 8), (55104, 512, 8)):
        run(nf, img, B)
