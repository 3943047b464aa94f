"""Fixed cost of one gom_conv3x3 launch: device time per call (graph-replayed) for layers of 1, 2, 4 ... items per tile on a
single 16 x 16 image (one CTA tile per 128 output channels), against the MMA work they contain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomavatar_b200 import conv as gconv  # noqa: E402


def time_it(fn, iters=20):
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


os.environ["GOM_CONV_SHAPE"] = "128,2,1"
for cin in (64, 128, 256, 512, 1024, 2048):
    for (n, S) in ((1, 16), (1, 32), (4, 64)):
        x = torch.randn(n, S, S, cin, device="cuda")
        w = torch.randn(128, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)
        wp = gconv.pack_weights(w)
        out = torch.empty(n, S, S, 128, device="cuda")
        t = time_it(lambda: gconv.conv3x3(x, wp, relu=True, out=out))
        tiles = n * (S // 16) ** 2
        mma_us = (cin // 32) * 9 * 8 * 64 / 1.9e3 * ((tiles + 147) // 148)
        print(f"c_in {cin:4d}  {n}x{S}x{S}: {tiles:3d} tiles, {cin // 32:2d} items/tile: {t:6.1f} us per call; MMA floor {mma_us:6.1f} us; overhead {t - mma_us:5.1f} us", flush=True)
