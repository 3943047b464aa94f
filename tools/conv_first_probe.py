"""Development tool: time csrc/conv_first.cu against cuDNN at the bench size (16 x 512 x 512 forward, 8 backward)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
import torch.nn.functional as F
from gomavatar_b200._lib import GomConvFirstArgs, call, ptr

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
    dev = "cuda:0"
    N, H, W = 16, 512, 512
    x = torch.randn(N, H, W, 3, device=dev)
    w = torch.randn(64, 3, 3, 3, device=dev) * 0.2
    b = torch.randn(64, device=dev)
    out = torch.empty(N, H, W, 64, device=dev)
    g = torch.randn(N // 2, H, W, 64, device=dev)
    dx = torch.empty(N // 2, H, W, 3, device=dev)
    f = lambda: call("gom_conv_first_forward", GomConvFirstArgs(n_images=N, height=H, width=W, x=ptr(x), weight=ptr(w), bias=ptr(b), out=ptr(out)))
    bw = lambda: call("gom_conv_first_backward", GomConvFirstArgs(n_images=N // 2, height=H, width=W, weight=ptr(w), dL_dout=ptr(g), dL_dx=ptr(dx)))
    xc = x.permute(0, 3, 1, 2)
    wc = w.contiguous(memory_format=torch.channels_last)
    gc = g.permute(0, 3, 1, 2)
    cf = lambda: torch.cudnn_convolution_relu(xc, wc, b, (1, 1), (1, 1), (1, 1), 1)
    cb = lambda: torch.ops.aten.convolution_backward(gc, xc[: N // 2], wc, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (True, False, False))
    print(f"own fwd {timeit(f):.3f} ms   cudnn fwd(+bias+relu) {timeit(cf):.3f} ms   own bwd {timeit(bw):.3f} ms   cudnn dgrad {timeit(cb):.3f} ms")
