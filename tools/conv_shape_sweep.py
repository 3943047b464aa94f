"""Time every tile shape / K-split of csrc/conv3x3_tc.cu on each VGG layer shape (GOM_CONV_SHAPE pins it), to calibrate the
cost model of gom_conv3x3.    python tools/conv_shape_sweep.py [--n-fwd 2]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomavatar_b200 import conv as gconv  # noqa: E402

LAYERS = [(64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 512, 64), (512, 512, 64), (512, 512, 32)]


def time_it(fn, iters=20):
    """device time per call: `iters` calls captured in ONE CUDA graph (host launch cost — ~30 us of Python + four tensor-map
    encodes per call — would otherwise bound the short layers), replayed 5 times"""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * iters) * 1e3


ap = argparse.ArgumentParser()
ap.add_argument("--n-fwd", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/conv_shape_sweep.json")
a = ap.parse_args()
res = {}
for (C, K, S) in LAYERS:
    for direction, n, cin, cout in (("fwd", a.n_fwd, C, K), ("dgrad", max(1, a.n_fwd // 2), K, C)):
        x = torch.randn(n, S, S, cin, device="cuda")
        w = torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)
        wp = gconv.pack_weights(w)
        out = torch.empty(n, S, S, cout, device="cuda")
        row = {}
        for shape in ["auto", "256,1,1"] + [f"{nt},{sub},{ks}" for nt in (128, 64) for sub in (2, 1) for ks in (1, 2, 4, 8)]:
            if shape != "auto":
                nt, sub, ks = map(int, shape.split(","))
                if cout % nt or (cin // 32) % ks or (cin // 32) // ks < 2:
                    continue
                os.environ["GOM_CONV_SHAPE"] = shape
            else:
                os.environ.pop("GOM_CONV_SHAPE", None)
            row[shape] = round(time_it(lambda: gconv.conv3x3(x, wp, relu=True, out=out)), 1)
        best = min((v, k) for k, v in row.items() if k != "auto")
        res[f"{cin}->{cout}@{S} x{n} {direction}"] = row
        print(f"{cin:3d}->{cout:3d}@{S:3d} x{n} {direction:5s} auto {row['auto']:6.1f} us | best {best[1]:8s} {best[0]:6.1f} us | " +
              " ".join(f"{k}:{v:.0f}" for k, v in row.items() if k != "auto"), flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
