"""Every VGG layer shape at a small image size: the automatically chosen tile shape / K-split / CTA pairing against a pinned
single shape, forward (bias + ReLU + mask) and dgrad (mask), 3xTF32 and TF32.  Prints the largest relative deviation per case."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomavatar_b200 import conv as C  # noqa: E402

LAYERS = [(64, 64, 128), (64, 128, 64), (128, 128, 64), (128, 256, 32), (256, 256, 32), (256, 512, 16), (512, 512, 16), (512, 512, 8)]
dev = "cuda:0"
torch.manual_seed(0)
for precision in ("fp32", "tf32"):
    for (ci, co, S) in LAYERS:
        for n in (2, 4, 8):
            x = torch.randn(n, S, S, ci, device=dev).relu()
            w = torch.randn(co, ci, 3, 3, device=dev) / (3 * ci ** 0.5)
            b = torch.randn(co, device=dev) * 0.1
            go = torch.randn(n, S, S, co, device=dev)
            res = {}
            for mode in ("auto", "pinned"):
                if mode == "pinned":
                    os.environ["GOM_CONV_SHAPE"] = "128,2" if co % 128 == 0 else "64,2"
                    os.environ["GOM_CONV_PAIR"] = "0"
                else:
                    os.environ.pop("GOM_CONV_SHAPE", None)
                    os.environ.pop("GOM_CONV_PAIR", None)
                st = torch.zeros(1, dtype=torch.int32, device=dev)
                m = C.new_mask(n, S, S, co, dev)
                y = C.conv3x3(x, C.pack_weights(w, split=precision == "fp32"), bias=b, relu=True, mask_out=m, precision=precision, status=st)
                mi = C.new_mask(n, S, S, ci, dev)
                xm = (x > 0)
                bits = (xm.reshape(n, S, S, ci // 32, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(-1)
                mi = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32).contiguous()
                g = C.conv3x3(go, C.pack_weights(w, transpose=True, split=precision == "fp32"), mask_in=mi, precision=precision, status=st)
                res[mode] = (y, g, int(st.item()))
            ey = float((res["auto"][0] - res["pinned"][0]).abs().max() / res["pinned"][0].abs().max())
            eg = float((res["auto"][1] - res["pinned"][1]).abs().max() / res["pinned"][1].abs().max())
            flag = "  <-----" if max(ey, eg) > (1e-5 if precision == "fp32" else 1e-9) else ""
            print(f"{precision} {ci:3d}->{co:3d}@{S:3d} x{n}: fwd {ey:.2e} dgrad {eg:.2e} status {res['auto'][2]}{flag}", flush=True)
