"""Layer-by-layer check of the tcgen05 shadow MLP on the GPU box (prints; exits non-zero on mismatch).

    timeout 300 python tools/shadow_debug.py

Compares every saved hidden activation and the output with float64 torch, first with probe weights that expose
operand-layout mistakes (each output feature copies ONE encoding column), then with random weights."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gomavatar_b200.shadow import FusedShadowModule          # noqa: E402
from gomavatar_b200.modules import posenc                     # noqa: E402


def run(m, x, label):
    m = m.cuda()
    xg = x.cuda().requires_grad_(True)                       # grad -> the kernel saves the hidden activations
    out = m(xg[None])[0, :, 0]
    torch.cuda.synchronize()
    ws = m._ws
    n_fg = int(ws["n_fg"].item())
    status = int(ws["status"].item())
    fg = (x != 0).any(dim=1).nonzero()[:, 0]
    print(f"[{label}] status {status}  n_fg {n_fg} (expected {fg.numel()})  bg {float(ws['bg_value'].item()):.7f}")
    idx_ok = torch.equal(ws["fg_index"][:n_fg].cpu().long(), fg) if n_fg == fg.numel() else False
    print(f"[{label}] fg_index ok: {idx_ok}")
    lin = [t for t in m.block_mlps if isinstance(t, torch.nn.Linear)]
    h = posenc(x[fg].double(), m.multires)
    bad = False
    for l, layer in enumerate(lin[:-1]):
        h = torch.relu(h @ layer.weight.detach().cpu().double().t() + layer.bias.detach().cpu().double())
        got = ws["hidden"][l, :, :n_fg].t().cpu().double()
        err = (got - h).abs().max().item() if n_fg else 0.0
        print(f"[{label}] hidden[{l}] max err {err:.3e}  (max |ref| {h.abs().max().item():.3e})")
        if err > 1e-4 * max(1.0, h.abs().max().item()):
            bad = True
            r, c = divmod(int((got - h).abs().argmax()), h.shape[1])
            print(f"   worst at row {r} feature {c}: got {got[r, c].item():.6f} ref {h[r, c].item():.6f}")
            print("   got[0, :8] ", np.round(got[0, :8].numpy(), 5))
            print("   ref[0, :8] ", np.round(h[0, :8].numpy(), 5))
            print("   got[:8, 0] ", np.round(got[:8, 0].numpy(), 5))
            print("   ref[:8, 0] ", np.round(h[:8, 0].numpy(), 5))
            break
    ref = torch.sigmoid(h @ lin[-1].weight.detach().cpu().double().t() + lin[-1].bias.detach().cpu().double())[:, 0]
    if not bad:
        err = (out.detach().cpu().double()[fg] - ref).abs().max().item() if n_fg else 0.0
        print(f"[{label}] out max err {err:.3e}")
        bad = err > 1e-5
    return bad or status != 0 or not idx_ok


def main():
    torch.manual_seed(0)
    bad = False
    cfg = {"multires": 6, "mlp_width": 128, "mlp_depth": 1, "skips": [9]}
    m = FusedShadowModule(cfg)
    with torch.no_grad():                                    # probe: feature n copies encoding column n % 39
        m.block_mlps[0].weight.zero_()
        m.block_mlps[0].bias.fill_(2.0)                      # keeps the ReLU open (|enc| <= 1.x for these inputs)
        for n in range(128):
            m.block_mlps[0].weight[n, n % 39] = 1.0
        m.block_mlps[-1].weight.fill_(0.01)
    x = torch.rand(300, 3) * 2 - 1
    x[::3] = 0
    bad |= run(m, x, "probe depth1")
    for depth in (1, 2, 3):
        cfg = {"multires": 6, "mlp_width": 128, "mlp_depth": depth, "skips": [9]}
        m = FusedShadowModule(cfg)
        with torch.no_grad():
            m.block_mlps[-1].weight.mul_(3e3)
        x = torch.randn(5000, 3)
        x[torch.rand(5000) < 0.4] = 0
        bad |= run(m, x, f"random depth{depth}")
    print("RESULT:", "MISMATCH" if bad else "OK")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
