"""One rasterizer case of tools/quick_raster_bench.py from the command line (what the ncu captures of profiles/r3_* ran):

    python tools/raster_case.py <n_gaussians> <image size> <iterations>      # 8 frames per launch
"""
import sys
sys.path.insert(0, "tools")
import quick_raster_bench as q
q.run(int(sys.argv[1]), int(sys.argv[2]), 8, iters=int(sys.argv[3]))
