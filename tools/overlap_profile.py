"""Development tool: do kernels of different streams overlap inside the captured training step?  torch.profiler (CUPTI) over
graph replays; prints the union busy time, the sum of kernel durations, and the time during which a convolution and a
non-convolution kernel ran simultaneously."""
import os, sys, json
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
import bench

if __name__ == "__main__":
    args = bench.parse()
    dev = torch.device("cuda:0")
    tr = bench.Trainer(args, 0, 1, dev)
    for i in range(4):
        tr.step_device(i)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(2):
            tr.step_device(4 + i)
        torch.cuda.synchronize()
    path = os.path.join(_R, "gpurun_out", "overlap_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    span = max(e["ts"] + e["dur"] for e in ev) - t0
    tot = sum(e["dur"] for e in ev)
    streams = sorted({e["args"].get("stream") for e in ev})
    # sweep line: busy union, conv-with-other overlap
    pts = []
    for e in ev:
        is_conv = "k_conv3x3" in e["name"]
        pts.append((e["ts"], 1, is_conv)); pts.append((e["ts"] + e["dur"], -1, is_conv))
    pts.sort()
    busy = both = 0.0
    n_conv = n_other = 0
    last = pts[0][0]
    for t, d, c in pts:
        if n_conv + n_other > 0: busy += t - last
        if n_conv > 0 and n_other > 0: both += t - last
        last = t
        if c: n_conv += d
        else: n_other += d
    print(f"streams {streams}; span {span/2e3:.3f} ms/step, sum of kernel durations {tot/2e3:.3f} ms/step, busy union {busy/2e3:.3f} ms/step, conv||other {both/2e3:.3f} ms/step")
    per = {}
    for e in ev:
        k = (e["args"].get("stream"), e["name"][:60])
        per[k] = per.get(k, 0) + e["dur"]
    for (s, n), d in sorted(per.items(), key=lambda x: -x[1])[:16]:
        print(f"  stream {s} {d/2e3:7.3f} ms  {n}")
    os.remove(path)
