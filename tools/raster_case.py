import sys
sys.path.insert(0, "tools")
import quick_raster_bench as q
q.run(int(sys.argv[1]), int(sys.argv[2]), 8, iters=int(sys.argv[3]))
