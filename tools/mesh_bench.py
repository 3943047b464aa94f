"""Development tool: csrc/mesh_raster.cu alone at the bench scene (30 000 faces, 512 x 512, sigma 1e-5, K = 50): forward /
backward times at B = 1 and B = 8 (CUDA events), and the new tile kernels against the one-block-per-tile kernel
(GOM_MESH_LEGACY=1)."""
import json
import math
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomavatar_b200 import synthetic as S  # noqa: E402
from gomavatar_b200.mesh_renderer import ndc_T_world, rasterize_mesh, vertex_normals  # noqa: E402
from gomavatar_b200.skinning import apply_lbs, get_global_RTs  # noqa: E402


def scene(n_faces, B, img, dev):
    sc = S.make_humanoid(n_faces, seed=0)
    pr = S.make_params(sc, seed=1)
    fr = S.make_frames(sc, B, img_size=(img, img), seed=100)
    t = lambda a: torch.from_numpy(a).to(dev)
    Rs, Ts = get_global_RTs(t(fr["cnl_gtfms"]), t(fr["dst_Rs"]), t(fr["dst_Ts"]))
    v = apply_lbs(t(pr["vertices"])[None].contiguous(), Rs, Ts, t(sc.lbs_weights))                     # [B,3,V]
    ndc = ndc_T_world(v, t(fr["K"]), t(fr["E"]), img, img).contiguous()
    vn = vertex_normals(v.permute(0, 2, 1).contiguous(), t(sc.faces).long())
    vn = torch.bmm(t(fr["E"])[:, :3, :3], vn.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
    return ndc, vn, t(sc.faces).long()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    out = {}
    blur = math.log(1. / 1e-4 - 1.) * 1e-5
    quick = "--quick" in sys.argv          # one configuration, new kernels only (for ncu)
    for n_faces, img in (((30000, 512),) if quick else ((30000, 512), (120000, 512))):
        for B in ((1,) if quick else (1, 8)):
            ndc, vn, faces = scene(n_faces, B, img, dev)
            cap = 16 * n_faces
            res = {}
            for legacy in (("0",) if quick else ("1", "0")):
                os.environ["GOM_MESH_LEGACY"] = legacy
                ndc_g, vn_g = ndc.clone().requires_grad_(True), vn.clone().requires_grad_(True)
                aux = {}
                nm, al, p2f = rasterize_mesh(ndc_g, vn_g, faces, img, img, soft=True, blur_radius=blur, faces_per_pixel=50, capacity=cap, aux=aux)
                g = torch.Generator(device="cpu").manual_seed(0)
                gn, ga = torch.randn(nm.shape, generator=g).to(dev), torch.randn(al.shape, generator=g).to(dev)
                (nm * gn).sum().backward(retain_graph=True) if False else ((nm * gn).sum() + (al * ga).sum()).backward()
                res[legacy] = dict(nm=nm.detach(), al=al.detach(), p2f=p2f, zcut=aux["zcut"].clone(), idcut=aux["idcut"].clone(),
                                   gv=ndc_g.grad.clone(), gn=vn_g.grad.clone(), status=int(aux["status"].max()))

                def fwd():
                    with torch.no_grad():
                        rasterize_mesh(ndc, vn, faces, img, img, soft=True, blur_radius=blur, faces_per_pixel=50, capacity=cap)

                def fwd_bwd():
                    a_, b_ = ndc.clone().requires_grad_(True), vn.clone().requires_grad_(True)
                    n_, l_, _ = rasterize_mesh(a_, b_, faces, img, img, soft=True, blur_radius=blur, faces_per_pixel=50, capacity=cap)
                    torch.autograd.backward([n_, l_], [gn, ga])
                res[legacy]["fwd_ms"] = timed(fwd)
                res[legacy]["fwd_bwd_ms"] = timed(fwd_bwd)
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(5):
                        fwd_bwd()
                    torch.cuda.synchronize()
                res[legacy]["kernels_us"] = {re.search(r"k_mesh_\w+", e.key).group(0): round(e.device_time_total / 5, 1)
                                             for e in prof.key_averages() if "k_mesh" in e.key}
            if quick:
                print(res["0"]["kernels_us"], res["0"]["fwd_ms"])
                return
            a, b = res["1"], res["0"]
            cut = torch.isfinite(a["zcut"])
            key = f"F{n_faces}_B{B}"
            out[key] = {
                "legacy_fwd_ms": a["fwd_ms"], "tiles_fwd_ms": b["fwd_ms"], "legacy_fwd_bwd_ms": a["fwd_bwd_ms"], "tiles_fwd_bwd_ms": b["fwd_bwd_ms"],
                "status": [a["status"], b["status"]], "legacy_kernels_us": a["kernels_us"], "tiles_kernels_us": b["kernels_us"],
                "pix_to_face_diff": float((a["p2f"] != b["p2f"]).float().mean()),
                "alpha_max_diff": float((a["al"] - b["al"]).abs().max()),
                "normal_max_diff": float((a["nm"] - b["nm"]).abs().max()),
                "pixels_with_cut": float(cut.float().mean()),
                "zcut_diff": float((a["zcut"][cut] != b["zcut"][cut]).float().mean()) if cut.any() else 0.0,
                "cut_set_diff": float((torch.isfinite(b["zcut"]) != cut).float().mean()),
                "idcut_diff": float((a["idcut"][cut] != b["idcut"][cut]).float().mean()) if cut.any() else 0.0,
                "grad_verts_rel": float((a["gv"] - b["gv"]).abs().max() / a["gv"].abs().max()),
                "grad_normals_rel": float((a["gn"] - b["gn"]).abs().max() / a["gn"].abs().max()),
            }
            print(key, json.dumps(out[key]), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mesh_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
