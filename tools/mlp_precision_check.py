"""Which of the two fp32 implementations of the non-rigid MLP (torch / cuBLAS fp32 vs the tcgen05 3xTF32 stack) is closer to a
float64 evaluation of the same module?  Prints max-relative errors of the outputs and the parameter gradients."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomavatar_b200 import modules as M  # noqa: E402

dev = "cuda:0"
cfg = {"name": "basic", "condition_code_size": 69, "mlp_width": 128, "mlp_depth": 6, "skips": [2], "multires": 6, "i_embed": 0,
       "kick_in_iter": 0, "full_band_iter": 10}
torch.manual_seed(5000)
net = M.NonRigidModule(cfg).to(dev)
with torch.no_grad():
    net.block_mlps[-1].weight.copy_(torch.randn_like(net.block_mlps[-1].weight) * 0.05)
    for m in net.block_mlps:
        if isinstance(m, torch.nn.Linear):
            m.bias.copy_(torch.randn_like(m.bias) * 0.1)
V, B = 5000, 3
xyz = (torch.randn(1, 3, V, device=dev) * 0.5)
pose = torch.randn(B, 69, device=dev) * 0.3
gout = torch.randn(B, 3, V, device=dev)


def run(module, x, p, g, tc):
    M._TC_MLP = tc
    module.zero_grad(set_to_none=True)
    x = x.clone().requires_grad_(True)
    out, _, _ = module(x, p, i_iter=1e7)
    (out * g).sum().backward()
    M._TC_MLP = True
    return out.detach(), x.grad, {n: q.grad.clone() for n, q in module.named_parameters()}


net64 = copy.deepcopy(net).double()
o64, gx64, gp64 = run(net64, xyz.double(), pose.double(), gout.double(), False)
rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
for name, tc in (("torch fp32", False), ("tcgen05 3xTF32", True)):
    o, gx, gp = run(net, xyz, pose, gout, tc)
    print(f"{name:16s} offsets {rel(o - xyz, o64 - xyz.double()):.2e}  dxyz {rel(gx, gx64):.2e}  params max "
          f"{max(rel(gp[k], gp64[k]) for k in gp):.2e}  (worst: {max(gp, key=lambda k: rel(gp[k], gp64[k]))})")
