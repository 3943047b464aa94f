"""How much of each tile's depth-sorted list does the blend actually consume?  (CPU oracle; planning data for the lazy /
bucketed per-tile sort queued in DESIGN.md §7.)  For every non-empty 16x16 tile: n = entries in its list, used = the largest
``n_contrib`` of its pixels (the deepest entry any pixel reached before it saturated or the list ended).

    python tools/sort_consumption.py 55104 512
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util_scene import raster_inputs  # noqa: E402
from oracle import raster as OR  # noqa: E402


def main(n_faces, img, frames=2):
    d = raster_inputs(n_faces=n_faces, img=img, n_frames=frames, channels=4)
    H, W = d["H"], d["W"]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    tot_n = tot_used = tot_sort = tot_sort_used = 0
    for b in range(frames):
        f = OR.forward(d["means3D"][b], d["cov6"][b], d["colors"], d["opacity"][b], d["view"][b], d["proj"][b],
                       float(d["tanfov"][b, 0]), float(d["tanfov"][b, 1]), d["bg"][b], H, W)
        ranges = f["ranges"].reshape(gy * gx, 2).astype(np.int64)
        n = ranges[:, 1] - ranges[:, 0]
        nc = f["n_contrib"].reshape(H, W)
        pad = np.zeros((gy * 16, gx * 16), nc.dtype)
        pad[:H, :W] = nc
        used = pad.reshape(gy, 16, gx, 16).max(axis=(1, 3)).reshape(-1).astype(np.int64)
        m = n > 0
        cost = lambda x: x * np.log2(np.maximum(x, 2)) ** 2          # bitonic network: n log^2 n
        tot_n += n[m].sum(); tot_used += used[m].sum(); tot_sort += cost(n[m]).sum(); tot_sort_used += cost(np.maximum(used[m], 1)).sum()
        print(f"frame {b}: tiles {m.sum()}, entries {n[m].sum()}, max list {n.max()}, consumed {used[m].sum()} "
              f"({used[m].sum() / n[m].sum():.1%}), median per-tile consumption {np.median(used[m] / n[m]):.1%}")
    print(f"n_faces {n_faces}, {img}^2: the blend reads {tot_used / tot_n:.1%} of the sorted entries; "
          f"ordering only that prefix would cost {tot_sort_used / tot_sort:.1%} of the full n log^2 n")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 30000, int(sys.argv[2]) if len(sys.argv) > 2 else 512)
