"""Layer-by-layer check of the tcgen05 shadow MLP (forward, saved images, backward) on the GPU box.

    timeout 300 python tools/shadow_debug.py

Decodes the activation / dZ images the kernels write and compares every one of them, the output and every gradient with
float64 torch; probe weights first (each output feature copies ONE encoding column: exposes operand-layout mistakes)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gomavatar_b200 import _lib                               # noqa: E402
from gomavatar_b200.shadow import FusedShadowModule          # noqa: E402
from gomavatar_b200.modules import posenc                     # noqa: E402

J, Q = np.meshgrid(np.arange(128), np.arange(32), indexing="ij")
SWZ = torch.from_numpy(J * 32 + ((((Q >> 2) ^ (J & 7)) << 2) | (Q & 3)))       # [128 features, 32 rows] -> word


def decode(words, base, n_feat, kb_stride):
    """[128 rows, n_feat] fp32 of one (tile, slot) whose row block kb starts at word base + kb * kb_stride"""
    rows = []
    for kb in range(4):
        off = base + kb * kb_stride
        v = words[off: off + n_feat * 32].view(torch.float32).double()[SWZ[:n_feat].reshape(-1)].reshape(n_feat, 32)
        rows.append(v.t())
    return torch.cat(rows, 0)


def report(label, got, ref, tol):
    scale = max(1e-30, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    bad = not (err <= tol * scale)
    print(f"[{label}] max err {err:.3e}  (max |ref| {scale:.3e}){'   <-- MISMATCH' if bad else ''}")
    if bad and got.dim() == 2:
        r, c = divmod(int((got - ref).abs().argmax()), ref.shape[1])
        print(f"   worst at ({r},{c}): got {got[r, c].item():.6g} ref {ref[r, c].item():.6g}")
        print("   got[0,:6]", np.round(got[0, :6].numpy(), 5), " ref[0,:6]", np.round(ref[0, :6].numpy(), 5))
        print("   got[:6,0]", np.round(got[:6, 0].numpy(), 5), " ref[:6,0]", np.round(ref[:6, 0].numpy(), 5))
    return bad


def run(m, x, label, tile=0):
    m = m.cuda()
    depth = len(m._linears()) - 1
    L = _lib.lib()
    act_words, dz_words = int(L.gom_shadow_mlp_tile_words(depth, 0)), int(L.gom_shadow_mlp_tile_words(depth, 1))
    xg = x.cuda().requires_grad_(True)
    g_out = torch.randn(x.shape[0], generator=torch.Generator().manual_seed(1))
    for p in m.parameters():
        p.grad = None
    out = m(xg[None])[0, :, 0]
    (out * g_out.cuda()).sum().backward()
    torch.cuda.synchronize()
    ws = m._ws
    n_fg, status = int(ws["n_fg"].item()), int(ws["status"].item())
    fg = (x != 0).any(dim=1).nonzero()[:, 0]
    print(f"[{label}] status {status}  n_fg {n_fg} (expected {fg.numel()})  capacity {ws['capacity']}")
    bad = status != 0 or n_fg != fg.numel() or not torch.equal(ws["fg_index"][:n_fg].cpu().long(), fg)
    # ---- float64 reference with autograd
    xd = x.double().requires_grad_(True)
    lin = m._linears()
    Ws = [l.weight.detach().cpu().double().requires_grad_(True) for l in lin]
    bs = [l.bias.detach().cpu().double().requires_grad_(True) for l in lin]
    h = posenc(xd, m.multires)
    acts, pre = [h], []
    for W, b in zip(Ws[:-1], bs[:-1]):
        z = h @ W.t() + b
        z.retain_grad()
        pre.append(z)
        h = torch.relu(z)
        acts.append(h)
    ref_out = torch.sigmoid(h @ Ws[-1].t() + bs[-1])[:, 0]
    (ref_out * g_out.double()).sum().backward()
    # ---- forward images, tile 0
    act = ws["act_img"].cpu()
    r0 = tile * 128
    n0 = min(128, n_fg - r0)
    rows = fg[r0: r0 + n0]
    for slot in range(depth + 1):
        nf = 64 if slot == 0 else 128
        base = tile * act_words + (0 if slot == 0 else 8192 + (slot - 1) * 16384)
        got = decode(act, base, nf, 2048 if slot == 0 else 4096)[:n0]
        ref = acts[slot].detach()[rows]
        if slot == 0:
            ref = torch.cat([ref, torch.zeros(n0, 64 - ref.shape[1], dtype=torch.float64)], 1)
        bad |= report(f"{label}] act slot {slot} tile {tile}", got, ref, 1e-5)
    bad |= report(f"{label}] out", out.detach().cpu().double(), ref_out.detach(), 1e-5)
    # ---- backward images, tile 0
    dz = ws["dz_img"].cpu()
    for l in range(depth):
        got = decode(dz, tile * dz_words + l * 16384, 128, 4096)[:n0]
        bad |= report(f"{label}] dZ layer {l} tile {tile}", got, pre[l].grad[rows], 1e-4)
    if tile > 0:                                            # every tile: which ones (and which quantity) go wrong first
        n_t = (n_fg + 127) // 128
        for name, words, tw, items in (("act", act, act_words, [(s_, 64 if s_ == 0 else 128, 0 if s_ == 0 else 8192 + (s_ - 1) * 16384,
                                                                  2048 if s_ == 0 else 4096) for s_ in range(depth + 1)]),
                                       ("dz", dz, dz_words, [(l, 128, l * 16384, 4096) for l in range(depth)])):
            for idx, nf, off, kbs in items:
                if name == "act":
                    ref_all = acts[idx].detach()[fg]
                    if idx == 0:
                        ref_all = torch.cat([ref_all, torch.zeros(ref_all.shape[0], 64 - ref_all.shape[1], dtype=torch.float64)], 1)
                else:
                    ref_all = pre[idx].grad[fg]
                scale = ref_all.abs().max().item()
                bad_tiles = []
                for t_ in range(n_t):
                    nn_ = min(128, n_fg - t_ * 128)
                    got = decode(words, t_ * tw + off, nf, kbs)[:nn_]
                    e_ = (got - ref_all[t_ * 128: t_ * 128 + nn_]).abs().max().item()
                    if e_ > 2e-5 * scale:
                        bad_tiles.append((t_, t_ % 148, t_ // 148, f"{e_ / scale:.1e}"))
                print(f"[{label}] {name} {idx}: {len(bad_tiles)} bad tiles of {n_t} (tile, cta, round, rel err): {bad_tiles[:12]}")
    gn_err = (xg.grad.cpu().double() - xd.grad).abs().max(dim=1).values
    worst = int(gn_err.argmax())
    pos = int((fg == worst).nonzero()[0, 0]) if bool((fg == worst).any()) else -1
    print(f"[{label}] worst g_normals pixel {worst}: fg row {pos} (tile {pos // 128 if pos >= 0 else -1}, row in tile {pos % 128 if pos >= 0 else -1}), "
          f"rows with err > 1e-3: {int((gn_err > 1e-3 * xd.grad.abs().max()).sum())}")
    n_kink = int((gn_err > 1e-3 * xd.grad.abs().max()).sum())           # a ReLU decided by the last bit moves ONE pixel's gradient
    report(f"{label}] g_normals", xg.grad.cpu().double(), xd.grad, 1e-3)
    bad |= n_kink > max(2, 1e-4 * x.shape[0]) or gn_err.max().item() > 5e-2 * xd.grad.abs().max().item()
    for i, l in enumerate(lin):
        bad |= report(f"{label}] grad W{i}", l.weight.grad.cpu().double(), Ws[i].grad, 1e-3)
        bad |= report(f"{label}] grad b{i}", l.bias.grad.cpu().double().reshape(-1, 1), bs[i].grad.reshape(-1, 1), 1e-3)
    return bad


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
        m.block_mlps[-1].weight.copy_(torch.linspace(-0.02, 0.02, 128)[None])
    x = torch.rand(300, 3) * 2 - 1
    x[::3] = 0
    bad |= run(m, x, "probe depth1")
    for depth in (1, 2, 3):
        cfg = {"multires": 6, "mlp_width": 128, "mlp_depth": depth, "skips": [9]}
        m = FusedShadowModule(cfg)
        with torch.no_grad():
            m.block_mlps[-1].weight.copy_(torch.randn(1, 128) * 0.15)
        x = torch.randn(5000, 3)
        x[torch.rand(5000) < 0.4] = 0
        bad |= run(m, x, f"random depth{depth}")
    m = FusedShadowModule({"multires": 6, "mlp_width": 128, "mlp_depth": 3, "skips": [9]})
    with torch.no_grad():
        m.block_mlps[-1].weight.copy_(torch.randn(1, 128) * 0.15)
    x = torch.randn(120_000, 3)
    x[torch.rand(120_000) < 0.5] = 0
    bad |= run(m, x, "multi-tile depth3", tile=300)
    print("RESULT:", "MISMATCH" if bad else "OK")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
