"""Development tool: eager vs CUDA-graph replay of the hot path (Model.forward + L1 losses + backward) per batch size."""
import os, sys, json
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import numpy as np, torch
from gomavatar_b200 import synthetic as S
from gomavatar_b200.losses import photometric_l1
from gomavatar_b200.model import Model, default_model_cfg

def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if __name__ == "__main__":
    dev = torch.device("cuda:0"); t = torch.from_numpy
    H = W = 512
    scene = S.make_humanoid(30000, seed=0)
    pr = S.make_params(scene, seed=1)
    work = torch.cuda.Stream()                 # everything off the legacy default stream (AccumulateGrad nodes keep their stream)
    torch.cuda.set_stream(work)
    for B in (1, 2, 8):
        m = Model(default_model_cfg(img_size=(W, H)), scene.canonical_info(), strict_raster=False).to(dev)
        with torch.no_grad():
            m.so3.copy_(t(pr["so3"])); m.scale.copy_(t(pr["scale"])); m.appearance_module.appearance.copy_(t(pr["appearance"]))
        fr = S.make_frames(scene, B, img_size=(W, H), seed=100)
        d = {k: t(v).to(dev) for k, v in fr.items()}
        gt = torch.rand(B, H, W, 3, device=dev); gtm = (torch.rand(B, H, W, device=dev) > 0.5).float()
        params = [p for p in m.parameters()]
        for p in params: p.grad = torch.zeros_like(p)
        def step():
            for p in params: p.grad.zero_()
            rgb, mask, _ = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"])
            _, a, b = photometric_l1(rgb, mask, d["bgcolor"], gt, gtm)
            (a + 5 * b).backward()
        eager = timeit(step)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=work):
            step()
        rep = timeit(graph.replay)
        print(json.dumps({"B": B, "eager_ms": eager, "graph_ms": rep, "eager_fps": B / eager * 1e3, "graph_fps": B / rep * 1e3}))
