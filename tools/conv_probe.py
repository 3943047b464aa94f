"""Development tool: time the VGG16 convolutions of the LPIPS trunk on the GPU (cuDNN through torch), per layer, for
the three ways the fused path can call them.  Decides LPIPS.conv_epilogue and the cudnn.benchmark setting."""
import os, sys, json
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
import torch.nn.functional as F

def timeit(fn, n=10):
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

if __name__ == "__main__":
    B2 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda:0"
    layers = [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128),
              (256, 512, 64), (512, 512, 64), (512, 512, 32)]
    for bench_mode in (False, True):
        torch.backends.cudnn.benchmark = bench_mode
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            tot = {"conv": 0.0, "conv_relu": 0.0, "dgrad": 0.0}
            mult = {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 2, 6: 1, 7: 2, 8: 3}
            for li, (ci, co, s) in enumerate(layers):
                x = torch.randn(B2, ci, s, s, device=dev).contiguous(memory_format=torch.channels_last)
                w = torch.randn(co, ci, 3, 3, device=dev).contiguous(memory_format=torch.channels_last) * 0.05
                b = torch.randn(co, device=dev)
                g = torch.randn(B2 // 2, co, s, s, device=dev).contiguous(memory_format=torch.channels_last)
                xi = x[: B2 // 2]
                t_conv = timeit(lambda: F.conv2d(x, w, None, padding=1))
                try:
                    t_cr = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1))
                except Exception as e:
                    t_cr = float("nan")
                t_dg = timeit(lambda: torch.ops.aten.convolution_backward(g, xi, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (True, False, False)))
                fl = 2 * 9 * ci * co * s * s * B2 / 1e12
                print(f"bench={bench_mode} tf32={tf32} {ci:4d}->{co:4d} @{s:3d}: conv {t_conv:7.3f} ms ({fl / t_conv * 1e3:6.0f} TF/s)  "
                      f"conv+bias+relu(cudnn) {t_cr:7.3f} ms  dgrad(B={B2 // 2}) {t_dg:7.3f} ms ({fl / 2 / t_dg * 1e3:6.0f} TF/s)", flush=True)
                tot["conv"] += t_conv * mult[li]; tot["conv_relu"] += t_cr * mult[li]; tot["dgrad"] += t_dg * mult[li]
            print(f"== bench={bench_mode} tf32={tf32} trunk totals (13 convs): {json.dumps(tot)}", flush=True)
