"""Kernel-level breakdown of the shadow MLP backward (torch profiler), 8x512^2, 20 % foreground."""
import sys
import torch
sys.path.insert(0, ".")
from gomavatar_b200.shadow import FusedShadowModule
torch.manual_seed(0)
N = 8 * 512 * 512
x = torch.randn(N, 3, device="cuda")
x[torch.rand(N, device="cuda") > 0.2] = 0
m = FusedShadowModule({"multires": 6, "mlp_width": 128, "mlp_depth": 3, "skips": [4]}, strict=False).cuda()
g = torch.randn(N, device="cuda")
def step():
    xg = x.detach().requires_grad_(True)
    out = m(xg[None])
    (out.reshape(-1) * g).sum().backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
