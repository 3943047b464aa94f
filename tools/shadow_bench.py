"""Microbenchmark of the shadow MLP on one B200: the fused tcgen05 path (gomavatar_b200.shadow.FusedShadowModule) against
the plain-torch module over every pixel (what the reference runs: fp32 cuBLAS, ~25 launches), forward and
forward+backward, on a batch of synthetic normal maps with a given foreground fraction.  Prints one JSON line.

    python tools/shadow_bench.py [--frames 8] [--img 512] [--fg 0.2] [--iters 20]
"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from gomavatar_b200 import _lib                              # noqa: E402
from gomavatar_b200.modules import ShadowModule              # noqa: E402
from gomavatar_b200.shadow import FusedShadowModule          # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.add_(1.0)                                       # > L2: evict between iterations
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--img", type=int, default=512)
    ap.add_argument("--fg", type=float, default=0.2)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    torch.manual_seed(0)
    dev = "cuda"
    N = args.frames * args.img * args.img
    # a disc of foreground per frame (compact like a body silhouette), normals |n| <= 3
    yy, xx = torch.meshgrid(torch.arange(args.img), torch.arange(args.img), indexing="ij")
    r2 = ((yy - args.img / 2) ** 2 + (xx - args.img / 2) ** 2).float()
    mask = (r2 <= args.fg * args.img * args.img / 3.14159265)[None].expand(args.frames, -1, -1).reshape(-1)
    x = torch.randn(N, 3)
    x = x / x.norm(dim=1, keepdim=True) * (1 + 2 * torch.rand(N, 1))
    x[~mask] = 0
    x = x.to(dev)
    cfg = {"multires": 6, "mlp_width": 128, "mlp_depth": 3, "skips": [4]}
    fused = FusedShadowModule(cfg, strict=False).to(dev)
    with torch.no_grad():
        fused.block_mlps[-1].weight.copy_(torch.randn(1, 128) * 0.15)
    ref = ShadowModule(cfg).to(dev)
    ref.load_state_dict(fused.state_dict())
    torch.backends.cuda.matmul.allow_tf32 = False             # the reference never enables TF32 matmuls
    flush = torch.zeros(256 << 20 >> 2, device=dev)
    g = torch.randn(N, device=dev)

    def fwd(m):
        with torch.no_grad():
            return m(x[None])

    def fwd_bwd(m):
        xg = x.detach().requires_grad_(True)
        out = m(xg[None])
        for p in m.parameters():
            p.grad = None
        (out.reshape(-1) * g).sum().backward()

    res = {"frames": args.frames, "img": args.img, "pixels": N, "n_fg": int(mask.sum()), "fg_frac": float(mask.float().mean())}
    _lib.profile_enable(True)
    res["fused_fwd_ms"] = timed(lambda: fwd(fused), args.iters, flush)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    for k in ("shadow_compact", "shadow_mlp_fwd"):
        if k in prof:
            res[k + "_ms"] = prof[k][0] / prof[k][1]
    _lib.profile_enable(True)
    res["fused_fwd_bwd_ms"] = timed(lambda: fwd_bwd(fused), args.iters, flush)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    for k in ("shadow_mlp_fwd", "shadow_mlp_bwd_data", "shadow_mlp_bwd_weights"):
        if k in prof:
            res["train_" + k + "_ms"] = prof[k][0] / prof[k][1]
    fused.strict = True
    fused.check_status()
    res["torch_fwd_ms"] = timed(lambda: fwd(ref), args.iters, flush)
    res["torch_fwd_bwd_ms"] = timed(lambda: fwd_bwd(ref), args.iters, flush)
    with torch.no_grad():
        res["max_abs_diff"] = float((fused(x[None]) - ref(x[None])).abs().max())
    flops = 2.0 * res["n_fg"] * (39 * 128 + 2 * 128 * 128 + 128)
    peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
    tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0        # TF32 runs at half the bf16 tensor rate
    res["alg_gflop_fwd"] = flops / 1e9
    res["mlp_kernel_alg_tflops"] = flops / (res.get("shadow_mlp_fwd_ms", res["fused_fwd_ms"]) * 1e-3) / 1e12
    res["mlp_kernel_issued_tflops"] = 3.0 * res["mlp_kernel_alg_tflops"]
    res["tf32_peak_tflops"] = tf32_peak
    res["tensor_frac_issued"] = res["mlp_kernel_issued_tflops"] / tf32_peak
    res["speedup_fwd"] = res["torch_fwd_ms"] / res["fused_fwd_ms"]
    res["speedup_fwd_bwd"] = res["torch_fwd_bwd_ms"] / res["fused_fwd_bwd_ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
