"""Development tool: where does the HOST time of an eager (no CUDA graph) training step go?  cProfile over a few steps of
bench.Trainer — the regime of the reference's unchanged train.py, which calls model / loss / backward / optimizer separately."""
import cProfile, os, pstats, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
import bench

if __name__ == "__main__":
    args = bench.parse()
    args.cuda_graph = False
    dev = torch.device("cuda:0")
    tr = bench.Trainer(args, 0, 1, dev)
    for i in range(5):
        tr.step_device(i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    n = 20
    for i in range(n):
        tr.step_device(5 + i)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime")
    print(f"total per step: {st.total_tt / n * 1e3:.2f} ms")
    st.print_stats(35)
    st.sort_stats("cumtime")
    st.print_stats(40)
