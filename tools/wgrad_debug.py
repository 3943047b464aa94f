"""Development tool: csrc/wgrad_tc.cu on small structured inputs (which operand element lands where)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomavatar_b200 import conv as C

dev = "cuda:0"
print("DEBUG", os.environ.get("GOM_WGRAD_DEBUG"))
torch.manual_seed(0)
for R, N in ((32, 32), (64, 128), (4096, 128), (4096, 192), (120272, 128)):
    g = torch.randn(R, 128, device=dev)
    x = torch.randn(R, N, device=dev)
    g_lo, x_lo = C.tf32_low_part(g), C.tf32_low_part(x)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(); t0 = time.time()
    gw = C.linear_wgrad(g, g_lo, x, x_lo, status=st)
    torch.cuda.synchronize(); dt = time.time() - t0
    ref = g.double().T @ x.double()
    print(R, N, "status", int(st[0]), f"{dt*1e3:.2f} ms", "max|gw|", float(gw.abs().max()), "err", float((gw.double() - ref).abs().max()), "ref max", float(ref.abs().max()))
    if R == 32:
        print(" gw[:4,:8]", gw[:4, :8].tolist())
        print(" g[:4,:8]", g[:4, :8].tolist())
        # one-hot probes: g = e_(r0, m0), x = e_(r0, n0) -> gw[m0, n0] = 1
        for (r0, m0, n0) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (9, 33, 5), (31, 127, 31)):
            g.zero_(); x.zero_(); g[r0, m0] = 1; x[r0, n0] = 1
            gw = C.linear_wgrad(g, torch.zeros_like(g), x, torch.zeros_like(x), status=st)
            nz = gw.nonzero().tolist()
            print("  probe", (r0, m0, n0), "->", nz[:6], [float(gw[i, j]) for i, j in nz[:6]])
