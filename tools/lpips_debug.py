"""Development tool: per-convolution gradient comparison, fused LPIPS path vs plain torch autograd (GPU, fp32)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import numpy as np, torch
import torch.nn as nn
from gomavatar_b200.lpips import LPIPS, seeded_random_trunk

dev = "cuda:0"
g = np.load(os.path.join(_R, "tests", "golden", "golden_lpips.npz"))
trunk = seeded_random_trunk(0)
heads = [g[f"lin{k}"] for k in range(5)]
t = torch.from_numpy
for name, (x0n, x1n) in {"golden": (g["x0"], g["x1"]),
                         "random": (np.random.default_rng(11).random((2, 3, 64, 48)).astype(np.float32),
                                    np.random.default_rng(12).random((2, 3, 64, 48)).astype(np.float32))}.items():
    # torch autograd reference with retained conv-output grads
    ref = LPIPS(trunk, heads, conv_precision="fp32", fused=False).to(dev)
    torch.backends.cudnn.allow_tf32 = False
    x0 = t(x0n).to(dev).requires_grad_(True)
    x1 = t(x1n).to(dev)
    conv_outs = []
    def taps(x):
        h = (x - ref.shift) / ref.scale
        h = h.contiguous(memory_format=torch.channels_last)
        outs = []
        for i, layer in enumerate(ref.features):
            h = layer(h)
            if isinstance(layer, nn.Conv2d) and h.requires_grad:
                h.retain_grad(); conv_outs.append(h)
            if i in (3, 8, 15, 22, 29):
                outs.append(h)
        return outs
    f0 = taps(2 * x0 - 1)
    with torch.no_grad():
        f1 = taps(2 * x1 - 1)
    total = 0
    for k, (a, b) in enumerate(zip(f0, f1)):
        d = (ref._unit(a) - ref._unit(b)) ** 2
        total = total + (d * getattr(ref, f"lin{k}")).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
    total.sum().backward()
    # fused, recording the gradient handed to each convolution's dgrad
    net = LPIPS(trunk, heads, conv_precision="fp32", fused=True).to(dev)
    rec = {}
    orig = net._conv_dgrad
    def spy(g_out, inp, ci):
        rec[ci] = g_out.detach().clone()
        return orig(g_out, inp, ci)
    net._conv_dgrad = spy
    k0 = t(x0n).to(dev).requires_grad_(True)
    val = net(2 * k0 - 1, 2 * x1 - 1)
    val.sum().backward()
    print(f"[{name}] value fused {val.reshape(-1).tolist()} torch {total.reshape(-1).tolist()}")
    for ci in range(13):
        a, b = rec[ci], conv_outs[ci].grad
        e = (a - b).abs()
        idx = np.unravel_index(int(e.argmax()), e.shape)
        print(f"[{name}] conv{ci:2d} {tuple(a.shape)} max|diff| {float(e.max()):.3e} / max|ref| {float(b.abs().max()):.3e} at {idx}; "
              f"n(|diff|>1e-3 max) {int((e > 1e-3 * b.abs().max()).sum())}")
    e = (k0.grad - x0.grad).abs()
    print(f"[{name}] input grad rel err {float(e.max() / x0.grad.abs().max()):.3e}")

# ---- near-tie check on the golden input: the level-1 (relu2_2) quad holding pixel (12, 2), channel 0
x1 = t(g["x1"]).to(dev); x0 = t(g["x0"]).to(dev)
torch.backends.cudnn.allow_tf32 = False
ref = LPIPS(trunk, heads, conv_precision="fp32", fused=False).to(dev)
with torch.no_grad():
    a_ref = ref._taps(2 * x0 - 1)[1]                          # torch path, batch of 2
    both = torch.cat([2 * x0 - 1, 2 * x1 - 1]).contiguous(memory_format=torch.channels_last)
    h = (both - ref.shift) / ref.scale
    h = h.contiguous(memory_format=torch.channels_last)
    for i, layer in enumerate(ref.features):
        h = layer(h)
        if i == 8:
            break
    a_cat = h[:2]                                              # same convolutions on the batch of 4
np.set_printoptions(precision=10)
print("quad (12:14, 2:4) ch0, torch batch-2 :", a_ref[0, 0, 12:14, 2:4].cpu().numpy().ravel())
print("quad (12:14, 2:4) ch0, batch-4       :", a_cat[0, 0, 12:14, 2:4].cpu().numpy().ravel())
print("max |a_ref - a_cat| level 1:", float((a_ref - a_cat).abs().max()))
