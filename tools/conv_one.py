"""Run each VGG layer shape of csrc/conv3x3_tc.cu once (forward, then dgrad) — the target of `ncu -k regex:k_conv3x3`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomavatar_b200 import conv as gconv  # noqa: E402

LAYERS = [(64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 512, 64), (512, 512, 64), (512, 512, 32)]
sel = [int(a) for a in sys.argv[1:]] or range(len(LAYERS))
for i in sel:
    C, K, S = LAYERS[i]
    torch.manual_seed(1)
    x = torch.randn(16, S, S, C, device="cuda").relu()
    w = torch.randn(K, C, 3, 3, device="cuda") / (3 * C ** 0.5)
    b = torch.randn(K, device="cuda") * 0.1
    mo = gconv.new_mask(16, S, S, K, "cuda")
    y = gconv.conv3x3(x, gconv.pack_weights(w), bias=b, relu=True, mask_out=mo)
    g = torch.randn(8, S, S, K, device="cuda")
    mi = gconv.new_mask(8, S, S, C, "cuda").fill_(-1)
    gx = gconv.conv3x3(g, gconv.pack_weights(w, transpose=True), mask_in=mi)
    torch.cuda.synchronize()
    print(i, LAYERS[i], float(y.sum()), float(gx.sum()))
