"""Development tool: per-kernel GPU time of one training step (torch.profiler / CUPTI)."""
import os, sys, json
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
import bench

if __name__ == "__main__":
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    args = bench.parse()
    dev = torch.device("cuda:0")
    tr = bench.Trainer(args, 0, 1, dev)
    for i in range(3):
        tr.step_device(i)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    n = 3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(n):
            tr.step_device(3 + i)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
        if e.device_type.name == "CUDA" or (t and e.key.startswith(("void", "sm", "cudnn", "k_", "ncclDev", "cutlass", "xmma", "implicit"))):
            rows.append((t / n / 1e3, e.count / n, e.key[:110]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"GPU kernel time per step: {tot:.2f} ms over {sum(r[1] for r in rows):.0f} launches")
    for ms, cnt, k in rows[:45]:
        print(f"{ms:8.3f} ms {ms / tot * 100:5.1f}%  x{cnt:6.1f}  {k}")
