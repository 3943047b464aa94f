"""Development tool: stage-by-stage parity of the B200 path against the CPU oracle at the BENCH size (512x512,
30 000 Gaussians).  Prints one JSON object: geometry error, rasterizer parity on identical inputs (bit-exact radii and
tile lists, image error statistics, PSNR), and the end-to-end Model.forward PSNR."""
import json, os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import numpy as np
import torch

from gomavatar_b200 import synthetic as S
from gomavatar_b200.model import Model, default_model_cfg
from gomavatar_b200.rasterizer import rasterize_gaussians
from gomavatar_b200.skinning import apply_lbs, face_gaussians, get_global_RTs
from oracle import camera as Cam, geometry as G, raster as R

t = torch.from_numpy


def stats(got, ref):
    err = np.abs(got - ref)
    mse = float((err.astype(np.float64) ** 2).mean())
    return {"max": float(err.max()), "frac_gt_1e-4rel": float((err > 1e-4 * np.abs(ref) + 1e-5).mean()),
            "n_gt_1e-3": int((err > 1e-3).sum()), "psnr_db": 999.0 if mse == 0 else float(-10 * np.log10(mse))}


if __name__ == "__main__":
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    img = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    dev = "cuda:0"
    H = W = img
    scene = S.make_humanoid(F, seed=0)
    pr = S.make_params(scene, seed=1)
    fr = S.make_frames(scene, 2, img_size=(W, H), seed=100)
    out = {"n_faces": scene.n_faces, "img": img}
    for b in range(2):
        _, xyz, cov = G.pose_geometry(t(pr["vertices"]), t(scene.faces), t(scene.lbs_weights), t(pr["so3"]), t(pr["scale"]),
                                      t(fr["cnl_gtfms"][b]), t(fr["dst_Rs"][b]), t(fr["dst_Ts"][b]))
        cov6 = G.pack_cov6(cov)
        st = Cam.raster_settings_from_KE(fr["K"][b], fr["E"][b], (W, H))
        app = pr["appearance"].T
        feat = np.ascontiguousarray(np.concatenate([app, np.ones_like(app[:, :1])], 1), dtype=np.float32)
        f = R.forward(xyz.numpy(), cov6.numpy(), feat, np.ones(scene.n_faces, np.float32), st.viewmatrix, st.projmatrix,
                      st.tanfovx, st.tanfovy, np.zeros(4, np.float32), H, W)
        ref = f["color"].transpose(1, 2, 0)
        # --- GPU geometry vs oracle geometry
        d = {k: t(fr[k][b:b + 1]).to(dev) for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts")}
        gR, gT = get_global_RTs(d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"])
        vo = apply_lbs(t(pr["vertices"]).to(dev)[None], gR, gT, t(scene.lbs_weights).to(dev))
        m3, c6 = face_gaussians(vo, t(scene.faces).to(dev), t(pr["so3"]).to(dev), t(pr["scale"]).to(dev), 1e-3)
        e_m = np.abs(m3[0].cpu().numpy() - xyz.numpy()).max()
        c6n, c6r = c6[0].cpu().numpy(), cov6.numpy()
        e_c = np.abs(c6n - c6r) / np.abs(c6r).max(axis=-1, keepdims=True)
        res = {"geom_mean_abs_err": float(e_m), "geom_cov_rel_err_max": float(e_c.max()),
               "geom_cov_rel_err_n_gt_1e-3": int((e_c.max(-1) > 1e-3).sum())}
        # --- rasterizer on IDENTICAL (oracle) geometry
        view, proj = t(st.viewmatrix)[None].to(dev), t(st.projmatrix)[None].to(dev)
        tanfov = torch.tensor([[st.tanfovx, st.tanfovy]], dtype=torch.float32, device=dev)
        aux = {}
        rgba, radii, _, _ = rasterize_gaussians(xyz[None].to(dev), cov6[None].to(dev), t(feat).to(dev), torch.ones(1, scene.n_faces, device=dev),
                                                view, proj, tanfov, torch.zeros(1, 4, device=dev), H, W, interleaved=True, aux=aux)
        res["raster_same_inputs"] = stats(rgba[0].cpu().numpy(), ref)
        res["radii_equal"] = bool(np.array_equal(radii[0].cpu().numpy(), f["radii"]))
        T = ((W + 15) // 16) * ((H + 15) // 16)
        off = aux["tile_offset"][0].cpu().numpy().astype(np.int64)
        res["n_dup"] = int(off[T])
        if "point_list" in f and "ranges" in f:
            pl = aux["point_list"][0].cpu().numpy()[: off[T]]
            res["point_list_equal"] = bool(np.array_equal(pl, np.asarray(f["point_list"])[: off[T]]))
        # --- rasterizer on GPU geometry, and the whole Model.forward
        rgba2, _, _, _ = rasterize_gaussians(m3, c6, t(feat).to(dev), torch.ones(1, scene.n_faces, device=dev), view, proj, tanfov,
                                             torch.zeros(1, 4, device=dev), H, W, interleaved=True)
        res["raster_gpu_geometry"] = stats(rgba2[0].cpu().numpy(), ref)
        m = Model(default_model_cfg(img_size=(W, H)), scene.canonical_info()).to(dev)
        with torch.no_grad():
            m.so3.copy_(t(pr["so3"])); m.scale.copy_(t(pr["scale"])); m.appearance_module.appearance.copy_(t(pr["appearance"]))
            rgb, mask, _ = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"])
        res["model_forward_rgb"] = stats(rgb[0].cpu().numpy(), ref[..., :3])
        res["model_forward_mask"] = stats(mask[0].cpu().numpy(), ref[..., 3])
        out[f"frame{b}"] = res
    print(json.dumps(out))
