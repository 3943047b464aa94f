#!/usr/bin/env python
"""bench.py — training-step throughput of the GoMAvatar hot path on B200 (contract: see the task brief / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200-native arm (this repo)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the reference's path via the oracle port

One "step" = forward (joint chain -> LBS -> face frame / covariance -> splat raster) + photometric losses (unpack, L1 rgb,
L1 mask, LPIPS-VGG) + backward + gradient all-reduce + Adam, over `--frames-per-step` frames per GPU at 512x512 with
30 000 Gaussians (BASELINE.json configs[2], synthetic stand-in for ZJU-MoCap 377: no dataset/checkpoint offline).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = None


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "train_step_frames_per_sec_512x512_30k_gaussians"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=8, help="frames per GPU per step (reference: 1)")
    ap.add_argument("--faces", type=int, default=30000)
    ap.add_argument("--img", type=int, default=512)
    ap.add_argument("--pool-steps", type=int, default=4, help="distinct batches cycled through")
    ap.add_argument("--lpips-precision", default="tf32", choices=["tf32", "fp32", "bf16"],
                    help="cuDNN conv precision of the LPIPS VGG trunk; tf32 = torch/cuDNN default = the reference's stock path")
    ap.add_argument("--lpips-torch", action="store_true", help="A/B: plain torch LPIPS glue instead of csrc/lpips.cu")
    ap.add_argument("--lpips-conv", default="tcgen05", choices=["tcgen05", "cudnn"],
                    help="VGG convolutions conv1_2..conv5_3: csrc/conv3x3_tc.cu (default, the product path) or the cuDNN A/B baseline")
    ap.add_argument("--lpips-streams", type=int, default=1,
                    help="groups of frames taken through the LPIPS network on separate CUDA streams (a group's HBM-bound tap kernels "
                         "overlap the other group's tensor-bound convolutions)")
    ap.add_argument("--side-stream", action="store_true",
                    help="(--full-model) run the mesh normal-map / shadow branch on a side stream next to the splat branch (no gain measured)")
    ap.add_argument("--lpips-epilogue", default="cudnn", choices=["kernel", "cudnn"],
                    help="(--lpips-conv cudnn only) bias+ReLU after each VGG convolution: own kernel, or cuDNN's fused conv-bias-activation")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="default: zero_grad + forward + losses + backward are captured once per input buffer set in a CUDA graph "
                         "and replayed (all-reduce and the Adam launch stay eager): +5 %% at 8 frames/step, +50 %% at 1 (launch-bound)")
    ap.add_argument("--full-model", action="store_true",
                    help="NOT the headline metric: the whole reference-shaped step of exps/zju-mocap_377.yaml — pose-refinement and "
                         "non-rigid MLPs, mesh normal map + soft silhouette (csrc/mesh_raster.cu), tcgen05 shadow MLP "
                         "(csrc/shadow_mlp.cu), rgb = albedo * shading, and the Laplacian / normal / colour regularisers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys b1 / full_model / strong_scaling")
    ap.add_argument("--cpu-frames", type=int, default=16, help="frames in the bounded CPU-baseline sample (~10 s of host work)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
# capture of this same command (profiles/); filled in by hand after each capture, None = not captured yet.
NCU_TRAFFIC_BYTES = {      # mean over the kernel's launches in one step, like `achieved`
    # profiles/r2_ncu_full_summary.md (8 frames, 512x512): tap_bwd 5 launches 3199.5 MB, tap_fwd 5 launches 2494.3 MB
    "lpips_tap_bwd": 639.9e6, "lpips_tap_fwd": 498.9e6,
    # profiles/r7_ncu_full_summary.md: conv1_1 forward = k_conv1_gemm<0> 1099.7 MB; backward = k_conv1_gemm<1> 629.2 MB +
    # k_conv1_stencil 118.4 MB; the CTA-pair convolutions: sum over the 12 forward (6 049 MB) / 12 dgrad (2 937 MB) launches of one step
    "conv_first_fwd": 1099.7e6, "conv_first_bwd": 747.6e6, "conv3x3_fwd": 504.1e6, "conv3x3_dgrad": 244.8e6}
# smsp__inst_executed.sum per launch (warp instructions) of the rasterizer's list kernels at the DEFAULT workload (8 frames,
# 30 000 Gaussians, 512x512, seeds of this file), from the committed ncu capture; None = not captured for this build.
NCU_WARP_INSTS = {"blend_fwd": 5.768e7, "blend_bwd": 1.053e8, "tile_sort": 2.507e7}
NCU_WARP_INSTS_SOURCE = "profiles/r5_ncu_full_summary.md (smsp__inst_executed.sum of k_blend<4>, k_blend_bwd<4,3,0>, k_tile_sort)"

_VGG_LEVELS = ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3))      # (channels, convolutions) per VGG16 block


def lpips_alg_bytes_per_frame(H, W, own_kernel_epilogue=False, fused_first_relu=True):
    """Algorithmic HBM bytes per FRAME (= one prediction + one target image) of the hand-written LPIPS kernels, summed
    over the layers each kernel serves (csrc/lpips.cu; formulas in DESIGN.md §4).  fp32 NHWC."""
    out = {"lpips_input": (2 * 2 + 2) * 3 * H * W * 4.0, "bias_relu": 0.0, "relu_bwd": 0.0, "lpips_tap_fwd": 0.0,
           "lpips_tap_bwd": 0.0,
           # conv1_1 in csrc/conv_first_tc.cu: forward for prediction + target; backward reads the gradient (and, with the
           # fused ReLU backward, the layer's own output) and writes 3 channels
           "conv_first_fwd": 2 * H * W * (3 + 64) * 4.0, "conv_first_bwd": H * W * (64 + 3 + (64 if fused_first_relu else 0)) * 4.0}
    h, w, cin = H, W, 3
    for level, (C, n_conv) in enumerate(_VGG_LEVELS):
        px = h * w
        if own_kernel_epilogue:
            out["bias_relu"] += n_conv * 2 * (2 * px * C * 4.0)          # read + write, pred and gt
        n_relu = n_conv - 1 - (1 if (level == 0 and fused_first_relu) else 0)   # conv1_1's ReLU backward lives in its dgrad kernel
        out["relu_bwd"] += n_relu * 3 * (px * C * 4.0)                   # act read, grad read + write (pred half)
        pooled = (h // 2) * (w // 2) * C * 4.0 if level < 4 else 0.0
        out["lpips_tap_fwd"] += 2 * px * C * 4.0 + 2 * pooled
        out["lpips_tap_bwd"] += 2 * px * C * 4.0 + pooled + px * C * 4.0
        h, w = h // 2, w // 2
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
class Trainer:
    def __init__(self, args, rank, world, device):
        from gomavatar_b200 import synthetic as S
        from gomavatar_b200.dist import FlatArena
        from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
        from gomavatar_b200.model import Model, default_model_cfg
        self.args, self.rank, self.world, self.dev = args, rank, world, device
        self.graphs, self.replays, self.graph_launches, self.graph_scope = {}, 0, 0, None
        if args.cuda_graph:          # everything off the legacy default stream: autograd's AccumulateGrad nodes keep the
            torch.cuda.set_stream(torch.cuda.Stream(device=device))     # stream they were created on, and capture needs one
        self.B = args.frames_per_step
        H = W = args.img
        scene = S.make_humanoid(args.faces, seed=0)
        self.scene = scene
        if args.full_model:           # the module nodes of reference exps/zju-mocap_377.yaml:63-98, all active (kick_in_iter 0)
            cfg = {"img_size": [W, H], "eval_mode": False,
                   "canonical_geometry": {"sigma": 1e-3, "radius_scale": 1.0, "deform_scale": True, "deform_so3": True},
                   "appearance": {"color_init": 0.5},
                   "non_rigid": {"name": "basic", "condition_code_size": 69, "mlp_width": 128, "mlp_depth": 6, "skips": [4],
                                 "multires": 6, "i_embed": 0, "kick_in_iter": 0, "full_band_iter": 50000},
                   "pose_refinement": {"name": "basic", "embedding_size": 69, "total_bones": 24, "mlp_width": 256, "mlp_depth": 4,
                                       "refine_root": False, "refine_t": False, "kick_in_iter": 0},
                   "normal_renderer": {"name": "mesh", "soft_mask": True, "sigma": 1e-5},
                   "shadow_module": {"name": "basic", "mlp_width": 128, "mlp_depth": 3, "skips": [4], "multires": 6, "i_embed": 0}}
            self.loss_cfg = {"rgb": {"coeff": 1.0}, "mask": {"coeff": 5.0}, "lpips": {"coeff": 1.0},       # configs/default.yaml:101-121
                             "laplacian": {"coeff_canonical": 0.0, "coeff_observation": 10.0},             # + exps/zju-mocap_377.yaml:101-112
                             "normal": {"mask_dilate": True, "kernel_size": 7, "coeff_mask": 1.0, "coeff_consist": 0.10},
                             "color_consist": {"coeff": 0.050}}
            self.model = Model(cfg, scene.canonical_info(), strict_raster=False).to(device)
            self.model.mesh_side_stream = bool(getattr(args, "side_stream", False))
            self.model.normal_renderer.strict = False             # overflow flags stay on the device (graph capture)
            self.model.normal_renderer.capacity = 16 * args.faces
            self.model.shadow_module.strict = False
            with torch.no_grad():                                 # the reference's 1e-5 last layers would make the MLPs invisible
                g = torch.Generator(device="cpu").manual_seed(5)
                self.model.shadow_module.block_mlps[-1].weight.copy_(torch.randn(1, 128, generator=g) * 0.15)
                self.model.non_rigid_module.block_mlps[-1].weight.copy_(torch.randn(3, 128, generator=g) * 2e-4)
                self.model.pose_refinement_module.block_mlps[-1].weight.copy_(torch.randn(69, 256, generator=g) * 2e-3)
        else:
            self.model = Model(default_model_cfg(img_size=(W, H)), scene.canonical_info(), strict_raster=False).to(device)
        pr = S.make_params(scene, seed=1)
        with torch.no_grad():
            self.model.so3.copy_(torch.from_numpy(pr["so3"]))
            self.model.scale.copy_(torch.from_numpy(pr["scale"]))
            self.model.appearance_module.appearance.copy_(torch.from_numpy(pr["appearance"]))
        self.model.train()
        heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
        self.lpips = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)], conv_precision=args.lpips_precision,
                           fused=not args.lpips_torch, conv_epilogue=args.lpips_epilogue, conv_impl=args.lpips_conv,
                           streams=getattr(args, "lpips_streams", 1)).to(device)
        # pool of frames: different poses / cameras / backgrounds per rank
        n_pool = self.B * args.pool_steps
        fr = S.make_frames(scene, n_pool, img_size=(W, H), seed=100 + rank)
        self.keys = ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor")
        self.host = {k: torch.from_numpy(fr[k]).pin_memory() for k in self.keys}
        self.devd = {k: v.to(device) for k, v in self.host.items()}
        # targets: render of a perturbed ("teacher") parameter set over the frame's random background
        self._make_targets(S, n_pool)
        self.arena = FlatArena(self.model)
        self.arena.broadcast_params()
        groups = self.model.get_param_groups(type("C", (), {"lr": {"appearance": 5e-4, "canonical_geometry": 5e-4,
                                                                "canonical_geometry_xyz": 5e-4, "non_rigid": 5e-4,
                                                                "pose_refinement": 5e-5, "shadow": 5e-4}})())
        from gomavatar_b200.dist import ArenaAdam
        # one launch over the flat arena (csrc/adam.cu); step counters on the device and the reference's exponential
        # learning-rate decay (train.py:166-175: lr * 0.1^(iter / 100 000)) evaluated there: nothing step-dependent is a
        # kernel argument, so the optimizer step is part of the captured graph
        self.opt = ArenaAdam(self.arena, groups, device_state=True, lr_decay=(0.1, 100000.0))
        self.h2d_bytes = sum(v[: self.B].numel() * v.element_size() for v in self.host.values()) + \
            self.host_tgt_rgb[: self.B].numel() * 4 + self.host_tgt_mask[: self.B].numel() * 4

    def _make_targets(self, S, n_pool):
        from gomavatar_b200.losses import unpack
        m = self.model
        rng = np.random.default_rng(7)
        saved = [p.detach().clone() for p in (m.vertices, m.appearance_module.appearance)]
        with torch.no_grad():
            m.vertices.add_(torch.from_numpy(rng.normal(0, 3e-3, tuple(m.vertices.shape)).astype(np.float32)).to(self.dev))
            m.appearance_module.appearance.copy_(torch.from_numpy(rng.uniform(0, 1, tuple(saved[1].shape)).astype(np.float32)).to(self.dev))
            rgbs, masks = [], []
            for s in range(0, n_pool, self.B):
                sl = slice(s, s + self.B)
                d = {k: v[sl] for k, v in self.devd.items()}
                rgb, mask, _ = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"], dst_posevec=d["dst_posevec"])
                rgbs.append(unpack(rgb, mask, d["bgcolor"]).clamp(0, 1)); masks.append(mask.clamp(0, 1).clone())
            m.vertices.copy_(saved[0]); m.appearance_module.appearance.copy_(saved[1])
        self.tgt_rgb, self.tgt_mask = torch.cat(rgbs).contiguous(), torch.cat(masks).contiguous()
        self.host_tgt_rgb, self.host_tgt_mask = self.tgt_rgb.cpu().pin_memory(), self.tgt_mask.cpu().pin_memory()

    def _fwd_bwd(self, d, tgt_rgb, tgt_mask, with_optimizer=False):
        from gomavatar_b200.losses import compute_loss
        self.arena.zero_grad()
        rgb, mask, outputs = self.model(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"], dst_posevec=d["dst_posevec"],
                                        bgcolor=d["bgcolor"])
        if self.args.full_model:
            from gomavatar_b200.regularizers import compute_loss as full_loss
            loss, _ = full_loss(rgb, mask, d["bgcolor"], tgt_rgb, tgt_mask, outputs, self.model, self.loss_cfg, lpips_func=self.lpips)
        else:
            loss, terms, _ = compute_loss(rgb, mask, d["bgcolor"], tgt_rgb, tgt_mask, lpips_func=self.lpips)
        loss.backward()
        if with_optimizer:
            self.opt.step(grad_scale=self.arena.all_reduce_sum())  # 1/world folded into the Adam launch
        return loss.detach()

    def _train(self, d, tgt_rgb, tgt_mask, graph_key=None):
        """graph_key: identity of a STATIC (d, tgt_rgb, tgt_mask) buffer set; with --cuda-graph the step is captured on first
        use and replayed afterwards.  Scope of the graph: "step" = zero_grad + forward + losses + backward + NCCL all-reduce +
        Adam (first choice); "fwd_bwd" = without the last two, which then stay eager (fallback if the collective cannot be
        captured); eager if that fails too."""
        if self.args.cuda_graph and graph_key is not None and graph_key not in self.graphs:
            from gomavatar_b200 import _lib
            _lib.profile_enable(False)                           # profiling events cannot be recorded inside a capture
            for scope in (("step", "fwd_bwd") if self.graph_scope is None else (self.graph_scope,)):
                try:
                    for _ in range(2):                           # warm-up on the capture stream (allocator, NCCL channels);
                        self._fwd_bwd(d, tgt_rgb, tgt_mask)      # no optimizer step: the parameters must not move here
                        self.arena.all_reduce_sum()
                    torch.cuda.synchronize(self.dev)
                    n0 = _lib.launch_count()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=torch.cuda.current_stream(self.dev)):
                        loss = self._fwd_bwd(d, tgt_rgb, tgt_mask, with_optimizer=scope == "step")
                    self.graphs[graph_key] = (g, loss)
                    self.graph_launches = _lib.launch_count() - n0
                    self.graph_scope = scope
                    break
                except Exception as e:
                    print(f"# CUDA-graph capture of scope '{scope}' failed ({type(e).__name__}: {e})", file=sys.stderr)
                    self.graph_error = f"{scope}: {type(e).__name__}: {e}"[:200]
                    torch.cuda.synchronize(self.dev)
            else:                                                # never lose the run to a capture problem: go eager
                self.args.cuda_graph = False
        if self.args.cuda_graph and graph_key is not None:
            g, loss = self.graphs[graph_key]
            g.replay()
            self.replays += 1
            if self.graph_scope != "step":
                self.opt.step(grad_scale=self.arena.all_reduce_sum())
        else:
            loss = self._fwd_bwd(d, tgt_rgb, tgt_mask, with_optimizer=True)
        return loss

    def step_device(self, i):
        s = (i % self.args.pool_steps) * self.B
        sl = slice(s, s + self.B)
        if self.args.cuda_graph:                                     # graphs read static buffers: D2D from the resident pool
            if not hasattr(self, "dev_static"):
                mk = lambda v: torch.empty((self.B,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.dev)
                self.dev_static = ({k: mk(v) for k, v in self.devd.items()}, mk(self.tgt_rgb), mk(self.tgt_mask))
            d, tr, tm = self.dev_static
            for k, v in self.devd.items():
                d[k].copy_(v[sl])
            tr.copy_(self.tgt_rgb[sl]); tm.copy_(self.tgt_mask[sl])
            return self._train(d, tr, tm, graph_key="dev")
        return self._train({k: v[sl] for k, v in self.devd.items()}, self.tgt_rgb[sl], self.tgt_mask[sl])

    # ---- end-to-end step: host buffers in, loss out, every step.  The input pipeline is the usual double-buffered one:
    # while step i computes, step i+1's batch (pinned host memory) is copied on a side stream into the other buffer
    # set, and the loss of step i-1 is read back (pinned, event-synchronised) so the host can run one step ahead.
    def _e2e_setup(self):
        if getattr(self, "_e2e_ready", False):
            return
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        mk = lambda v: torch.empty((self.B,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.dev)
        self.e2e_bufs = [({k: mk(v) for k, v in self.host.items()}, mk(self.host_tgt_rgb), mk(self.host_tgt_mask)) for _ in range(2)]
        self.e2e_ready_ev = [torch.cuda.Event(), torch.cuda.Event()]           # batch has landed in buffer set s
        self.e2e_done_ev = [None, None]                                        # compute that used buffer set s has finished
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_ev = [None, None]
        self.e2e_prefetched = None
        self.losses_read = 0
        self._e2e_ready = True

    def _prefetch(self, i):
        s = i % 2
        st = (i % self.args.pool_steps) * self.B
        sl = slice(st, st + self.B)
        d, tr, tm = self.e2e_bufs[s]
        with torch.cuda.stream(self.copy_stream):
            if self.e2e_done_ev[s] is not None:
                self.copy_stream.wait_event(self.e2e_done_ev[s])               # do not overwrite a batch still in use
            for k, v in self.host.items():
                d[k].copy_(v[sl], non_blocking=True)
            tr.copy_(self.host_tgt_rgb[sl], non_blocking=True)
            tm.copy_(self.host_tgt_mask[sl], non_blocking=True)
            self.e2e_ready_ev[s].record(self.copy_stream)
        self.e2e_prefetched = i

    def step_e2e(self, i):
        self._e2e_setup()
        s = i % 2
        if self.e2e_prefetched != i:
            self._prefetch(i)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.e2e_ready_ev[s])
        self._prefetch(i + 1)                                                  # overlaps with this step's compute
        d, tr, tm = self.e2e_bufs[s]
        loss = self._train(d, tr, tm, graph_key=("e2e", s))
        self.loss_host[s].copy_(loss.detach(), non_blocking=True)              # D2H of this step's loss
        ev = torch.cuda.Event(); ev.record(cur)
        self.loss_ev[s], self.e2e_done_ev[s] = ev, ev
        p = 1 - s
        if self.loss_ev[p] is not None:                                        # read the previous step's loss on the host
            self.loss_ev[p].synchronize()
            self.last_loss = float(self.loss_host[p]); self.loss_ev[p] = None; self.losses_read += 1

    def flush_e2e(self):
        for s in range(2):
            if self.loss_ev[s] is not None:
                self.loss_ev[s].synchronize()
                self.last_loss = float(self.loss_host[s]); self.loss_ev[s] = None; self.losses_read += 1


def timed_region(fn, steps, warmup, world, device, flush=None):
    """W warm-ups, barrier + sync, K steps between CUDA events on the launching stream, barrier + sync, max over ranks.
    `flush` (host-side completion of pipelined result reads) runs inside the timed region."""
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    if flush is not None:
        flush()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    if flush is not None:
        flush()
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def vgg_conv_flops(H, W, n_images, skip_first=True):
    """2 * MACs of the VGG16 convolutions conv1_2 ... conv5_3 (conv1_1 too unless skip_first) over n_images H x W images"""
    total, h, w, cin, first = 0.0, H, W, 3, True
    for C, n_conv in _VGG_LEVELS:
        for _ in range(n_conv):
            if not (first and skip_first):
                total += 2.0 * 9 * cin * C * h * w * n_images
            first, cin = False, C
        h, w = h // 2, w // 2
    return total


def library_share(tr, steps=2):
    """Which kernels the step runs, by owner, from a short torch.profiler (CUPTI) trace of EAGER steps on rank 0: this repo's
    kernels (names from libgom_b200.so) / NCCL / memcpy+memset / anything else (torch ATen, cuDNN, cuBLAS, CUTLASS =
    "library").  Shares of the summed kernel time; the trace itself is not part of any timed region."""
    try:
        from torch.profiler import ProfilerActivity, profile
        was = tr.args.cuda_graph
        tr.args.cuda_graph = False
        tr.step_device(0)
        torch.cuda.synchronize(tr.dev)
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(steps):
                tr.step_device(i)
            torch.cuda.synchronize(tr.dev)
        tr.args.cuda_graph = was
        own = lib = nccl = mem = 0.0
        lib_names = {}
        for e in prof.events():
            if "cuda" not in str(getattr(e, "device_type", "")).lower():
                continue
            t = float(getattr(e, "device_time_total", 0.0) or getattr(e, "cuda_time_total", 0.0) or 0.0)
            n = e.name
            if re.search(r"(^|[\s:])k_[a-z0-9_]+", n):           # every kernel of libgom_b200.so is named k_*
                own += t
            elif "nccl" in n.lower():
                nccl += t
            elif n.lower().startswith("memcpy") or n.lower().startswith("memset"):
                mem += t
            else:
                lib += t
                lib_names[n[:60]] = lib_names.get(n[:60], 0.0) + t
        tot = own + lib + nccl + mem
        if tot <= 0:
            return {"unavailable": "the profiler returned no device events"}
        top = sorted(lib_names.items(), key=lambda kv: -kv[1])[:5]
        return {"own_kernels": own / tot, "library_kernels": lib / tot, "nccl": nccl / tot, "memcpy_memset": mem / tot,
                "device_us_per_step": tot / steps, "top_library_kernels": [[k, v / steps] for k, v in top],
                "how": "torch.profiler CUDA activity over %d eager steps, shares of summed device time" % steps}
    except Exception as e:                                    # CUPTI may be unavailable: report, never fail the bench
        return {"unavailable": f"{type(e).__name__}: {e}"[:160]}


def measure(args, rank, local, world, device, detailed):
    """Build the trainer for `args`, time the e2e and the device-resident step; with `detailed` also the per-kernel eager
    region, the rooflines and the clock samples."""
    from gomavatar_b200 import _lib
    tr = Trainer(args, rank, world, device)
    B, K, W_ = args.frames_per_step, args.steps, max(args.warmup, 3)
    frames_total = K * B * world
    ms_e2e = timed_region(tr.step_e2e, K, W_, world, device, flush=tr.flush_e2e)
    assert tr.losses_read == K + W_, "every e2e step must deliver its loss to the host"
    sampler = ClockSampler(local)
    if rank == 0 and detailed:
        sampler.start()
    r0, n0 = tr.replays, _lib.launch_count()
    ms = timed_region(tr.step_device, K, W_, world, device)
    used_graph = bool(args.cuda_graph)                           # False if the capture fell back to eager
    if used_graph:
        per_replay = tr.graph_launches + (0 if tr.graph_scope == "step" else 1)
        launches_timed = int(round((tr.replays - r0) * K / (K + W_))) * per_replay
    else:
        launches_timed = int(round((_lib.launch_count() - n0) * K / (K + W_)))
    clocks = sampler.stop() if (rank == 0 and detailed) else None
    out = {"value": frames_total / (ms * 1e-3), "ms_per_step": ms / K, "e2e_value": frames_total / (ms_e2e * 1e-3),
           "e2e_ms_per_step": ms_e2e / K, "gpu_launches": launches_timed, "clocks": clocks, "h2d": int(tr.h2d_bytes),
           "cuda_graph": (tr.graph_scope if used_graph else (getattr(tr, "graph_error", None) or False)), "trainer": tr}
    if not detailed:
        return out
    args.cuda_graph = False                                      # per-kernel CUDA-event timers need eager launches:
    _lib.profile_enable(True)                                    # a second, eager region feeds `kernels` / the rooflines
    ms_prof = timed_region(tr.step_device, K, W_, world, device)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    args.cuda_graph = used_graph
    out.update({"prof": prof, "ms_prof": ms_prof})
    return out


def run_b200(args):
    import copy
    from gomavatar_b200.dist import init_from_env
    rank, local, world = init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    B, K, W_ = args.frames_per_step, args.steps, max(args.warmup, 3)
    m = measure(args, rank, local, world, device, detailed=True)
    tr, prof, ms_prof, ms = m["trainer"], m["prof"], m["ms_prof"], m["ms_per_step"] * K

    # ---- per-kernel table (eager region, CUDA events on the launching stream) and the rooflines
    aux = tr.model.last_raster_aux
    T = ((args.img + 15) // 16) ** 2
    n_dup = float(aux["tile_offset"][:, T].to(torch.int64).bitwise_and(0xFFFFFFFF).float().mean().item())
    overflow = int(aux["status"].max().item())
    HW, F, V = args.img * args.img, tr.scene.n_faces, tr.scene.n_vertices
    alg_bytes_per_frame = {      # SURVEY.md §8d algorithmic bytes per frame (DESIGN.md §4 lists every formula)
        "blend_fwd": 40 * n_dup + 24 * HW, "blend_bwd": 80 * n_dup + 44 * HW, "tile_sort": 12 * n_dup, "preprocess": 76 * F,
        "preprocess_bwd": 76 * F + 36 * F, "emit": 12 * n_dup + 20 * F, "scan_tiles": 12 * T, "worklist": 12 * T,
        "lbs_fwd": 120 * V, "lbs_bwd": 120 * V, "face_fwd": 72 * F, "face_bwd": 72 * F + 36 * F,
        "photo_fwd": 44 * HW, "photo_bwd": 60 * HW}
    # tcgen05 path: conv1_1's ReLU backward is applied by conv1_2's dgrad through the bit mask the forward emits, so the backward of
    # conv1_1 does not read its own activation
    alg_bytes_per_frame.update(lpips_alg_bytes_per_frame(args.img, args.img, args.lpips_conv == "cudnn" and args.lpips_epilogue == "kernel",
                                                         fused_first_relu=args.lpips_conv != "tcgen05"))
    if args.lpips_conv == "tcgen05":
        alg_bytes_per_frame["relu_bwd"] = 0.0                            # fused into the dgrad epilogues (bit masks)
    alg_bytes_per_frame["adam"] = 28.0 * tr.arena.numel / B          # param r/w, grad r, two moments r/w: per STEP
    alg_flops_per_frame = {"conv3x3_fwd": vgg_conv_flops(args.img, args.img, 2), "conv3x3_dgrad": vgg_conv_flops(args.img, args.img, 1)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0      # TF32 runs at half the bf16 tensor rate
    tpeak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 = dense TF32 (of measured)" if "bf16_tflops_sustained" in peaks
                 else "1400 / 2 TFLOP/s (of fallback)")
    kernels = {}
    n_timed_steps = K + W_
    for name, (tot_ms, n) in prof.items():
        ms_step = tot_ms / n_timed_steps                     # all launches of this kernel in one step
        per_step = n / n_timed_steps
        k = {"ms_per_step": ms_step, "launches_per_step": per_step, "ms_per_launch": tot_ms / max(n, 1),
             "share_of_step": ms_step / (ms_prof / K)}
        if name in alg_flops_per_frame:
            fl = alg_flops_per_frame[name] * B
            k.update({"alg_flops_per_launch": fl / max(per_step, 1e-9), "tflops": fl / (ms_step * 1e-3) / 1e12 if ms_step > 0 else None})
        else:
            byt_step = alg_bytes_per_frame.get(name, 0.0) * B    # algorithmic bytes of all those launches
            k.update({"alg_bytes_per_launch": byt_step / max(per_step, 1e-9), "gbs": byt_step / (ms_step * 1e-3) / 1e9 if ms_step > 0 else None})
        kernels[name] = k
    bounded = [k for k in kernels if alg_bytes_per_frame.get(k, 0.0) > 0 or k in alg_flops_per_frame]
    dom = max(bounded, key=lambda k: kernels[k]["ms_per_step"]) if bounded else None
    roofline = None
    if dom:
        k = kernels[dom]
        if dom in alg_flops_per_frame:
            roofline = {"kernel": dom, "bound": "tensor", "achieved": k["tflops"], "peak": tpeak, "unit": "TFLOP/s",
                        "frac": k["tflops"] / tpeak, "traffic": NCU_TRAFFIC_BYTES.get(dom), "peak_source": tpeak_src, "ms_per_launch": k["ms_per_launch"],
                        "alg_flops_per_launch": k["alg_flops_per_launch"], "launches_per_step": k["launches_per_step"],
                        "note": "TF32 tcgen05 implicit-GEMM convolutions on CTA pairs (cta_group::2; 12 VGG layers per step, mean over the "
                                "launches); achieved = 2 * MACs / CUDA-event time; frac > 1: the peak is half of cuBLAS's SUSTAINED bf16 "
                                "rate (power-limited, ~1.4 GHz), which these TF32 kernels exceed at the step's duty cycle",
                        "frac_of_burst": k["tflops"] / (float(peaks.get("bf16_tflops", 0.0)) / 2.0) if peaks.get("bf16_tflops") else None}
        else:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
                        "traffic": NCU_TRAFFIC_BYTES.get(dom), "peak_source": peak_src, "ms_per_launch": k["ms_per_launch"],
                        "alg_bytes_per_launch": k["alg_bytes_per_launch"], "launches_per_step": k["launches_per_step"]}
    # the rasterizer's list kernels: the north star's ">= 60 % of HBM" target next to the bound that actually binds them
    sm_mhz = (m["clocks"] or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    default_workload = (args.faces == 30000 and args.img == 512 and B == 8 and not args.full_model)
    roofline_blend = {"n_dup_per_frame": n_dup, "peak_hbm_gbs": peak, "peak_source": peak_src,
                      "issue_peak": "148 SMs x 4 schedulers x SM clock (median under load: %.0f MHz) warp instructions / s" % sm_mhz,
                      "warp_insts_source": NCU_WARP_INSTS_SOURCE if default_workload else None, "kernels": {}}
    for kk in ("tile_sort", "blend_fwd", "blend_bwd"):
        if kk in kernels:
            e = {"ms_per_launch": kernels[kk]["ms_per_launch"], "alg_bytes_per_launch": kernels[kk]["alg_bytes_per_launch"],
                 "gbs": kernels[kk]["gbs"], "frac_hbm": kernels[kk]["gbs"] / peak}
            wi = NCU_WARP_INSTS.get(kk) if default_workload else None
            if wi:
                e["warp_insts_per_launch"] = wi
                e["frac_issue"] = wi / (148 * 4 * sm_mhz * 1e6 * kernels[kk]["ms_per_launch"] * 1e-3)
            roofline_blend["kernels"][kk] = e
    roofline_conv = None
    if "conv3x3_fwd" in kernels:
        roofline_conv = {"bound": "tensor", "unit": "TFLOP/s", "peak": tpeak, "peak_source": tpeak_src,
                         "kernels": {kk: {"tflops": kernels[kk]["tflops"], "frac": kernels[kk]["tflops"] / tpeak,
                                          "ms_per_step": kernels[kk]["ms_per_step"], "launches_per_step": kernels[kk]["launches_per_step"]}
                                     for kk in ("conv3x3_fwd", "conv3x3_dgrad") if kk in kernels}}

    if args.full_model and roofline is not None and "shadow_mlp_fwd" in kernels:
        sm = tr.model.shadow_module
        n_fg = int(sm._ws["n_fg"].item())
        flops = 2.0 * n_fg * (39 * 128 + 2 * 128 * 128 + 128)           # one pass over the MLP (fp32-equivalent FLOPs)
        roofline["shadow_mlp"] = {"bound": "tensor", "unit": "TFLOP/s", "peak": tpeak, "n_fg_per_step": n_fg,
                                  "peak_source": tpeak_src,
                                  "note": "achieved = algorithmic fp32-equivalent FLOPs / time; every product is issued as 3 TF32 MMAs"}
        for kk, mult in (("shadow_mlp_fwd", 1.0), ("shadow_mlp_bwd_data", 1.0), ("shadow_mlp_bwd_weights", 1.0)):
            if kk in kernels:
                ach = flops * mult / (kernels[kk]["ms_per_launch"] * 1e-3) / 1e12
                roofline["shadow_mlp"][kk] = {"achieved": ach, "frac": ach / tpeak, "frac_issued": 3 * ach / tpeak,
                                              "ms_per_launch": kernels[kk]["ms_per_launch"]}
    lib = library_share(tr) if (rank == 0 and world == 1) else None

    # ---- extra measurements of the same step (extra keys; the headline above is untouched)
    extras = {}
    del m["trainer"], tr
    torch.cuda.empty_cache()

    def extra(**kw):
        a2 = copy.copy(args)
        for k_, v_ in kw.items():
            setattr(a2, k_, v_)
        a2.steps, a2.warmup = max(5, K // 2), 3
        r = measure(a2, rank, local, world, device, detailed=False)
        del r["trainer"]
        torch.cuda.empty_cache()
        return {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "e2e_value": r["e2e_value"],
                "frames_per_step_per_gpu": a2.frames_per_step, "global_batch": a2.frames_per_step * world, "steps": a2.steps,
                "cuda_graph": r["cuda_graph"], "gpu_launches": r["gpu_launches"]}
    if not args.no_extras and not args.full_model:
        if world == 1:
            if B != 1:        # the reference's own batch size (configs/default.yaml:10, gaussian.py:24): one Adam step per frame
                extras["b1"] = extra(frames_per_step=1)
            extras["full_model"] = extra(full_model=True)
            extras["full_model"]["note"] = "whole reference-shaped step of exps/zju-mocap_377.yaml (bench.py --full-model)"
            if B != 1:        # ... and exactly what the reference's train.py runs: the full model at one frame per optimizer step
                extras["full_model_b1"] = extra(full_model=True, frames_per_step=1)
        elif B % world == 0:  # fixed global batch of B frames split over the ranks (the headline keeps B frames per GPU)
            extras["strong_scaling"] = extra(frames_per_step=B // world)

    line = None
    if rank == 0:
        line = {
            "metric": ("full_model_" + METRIC) if args.full_model else METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (every kernel hand-written; VGG convolutions of LPIPS: %s products on tcgen05, fp32 accumulation)"
                     % ("TF32" if args.lpips_precision == "tf32" else args.lpips_precision) if args.lpips_conv == "tcgen05" else
                     "f32 (LPIPS VGG convs in cuDNN at %s: A/B baseline, not the product path)" % args.lpips_precision,
            "data": "synthetic (seeded SMPL-topology humanoid, poses, ZJU-like cameras; LPIPS trunk = seeded random VGG16, "
                    "no ImageNet weights offline)",
            "config": {"workload": ("FULL reference-shaped train step of exps/zju-mocap_377.yaml (hot path + pose-refinement / non-rigid "
                                    "MLPs + mesh normal map + tcgen05 shadow MLP + Laplacian / normal / colour regularisers + Adam), "
                                    "512x512, 30k Gaussians" if args.full_model else
                                    "ZJU-MoCap-377-like train step (LBS+face frame+splat raster+L1/LPIPS losses+backward+Adam), "
                                    "512x512, 30k Gaussians (BASELINE configs[2])"),
                       "img": args.img, "n_gaussians": F, "n_vertices": V, "frames_per_step_per_gpu": B,
                       "global_batch": B * world, "parallelism": f"frame-sharded dp{world}, 1 NCCL all-reduce of the flat grad arena/step",
                       "lpips_conv": args.lpips_conv, "lpips_conv_precision": args.lpips_precision + (" (the reference's stock cuDNN setting: it never disables TF32)" if args.lpips_precision == "tf32" else ""),
                       "raster_overflow": overflow, "cuda_graph": m["cuda_graph"],
                       "l2": "per-step working set (LPIPS activations, ~%d MB) exceeds the 126 MB L2; batches cycle through a pool"
                             % int(B * 2 * 32e6 * 4 / 1e6)},
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "ms_per_step": m["e2e_ms_per_step"], "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": 4},
            "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "roofline": roofline, "roofline_blend": roofline_blend,
            "roofline_conv": roofline_conv, "library_share_of_step": lib, "kernels": kernels,
        }
        line.update(extras)
        if not args.no_cpu_baseline and world == 1 and not args.full_model:
            line["cpu_baseline"] = cpu_path(args, n_frames=args.cpu_frames, gpu_device=device)
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return line


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_path(args, n_frames, gpu_device=None, steps=1, warmup=0):
    """The reference's path on the host cores: its PyTorch ops restated (oracle/geometry.py, losses.py) + the C
    restatement of the un-vendored CUDA rasterizer (oracle/raster_oracle.c, OpenMP), forward + backward, n_frames frames
    per step, batch 1 per frame like the reference.  The ONLY place bench.py touches oracle/."""
    from gomavatar_b200 import synthetic as S
    from oracle import camera as Cam, geometry as G, losses as OL, raster as R
    H = W = args.img
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t = torch.from_numpy
    scene = S.make_humanoid(args.faces, seed=0)
    pr = S.make_params(scene, seed=1)
    fr = S.make_frames(scene, max(n_frames, 1), img_size=(W, H), seed=100)
    heads = np.load(os.path.join(ROOT, "tests", "golden", "golden_lpips.npz"))
    lp = OL.LPIPSVGG(OL.seeded_random_trunk_state(0), [heads[f"lin{k}"] for k in range(5)])
    rng = np.random.default_rng(3)
    tgt = t(rng.random((H, W, 3)).astype(np.float32))[None]
    tgt_m = t((rng.random((H, W)) > 0.5).astype(np.float32))[None]
    psnr = None

    def one_frame(b):
        nonlocal psnr
        v = t(pr["vertices"]).requires_grad_(True)
        so3, sc = t(pr["so3"]).requires_grad_(True), t(pr["scale"]).requires_grad_(True)
        app = t(pr["appearance"]).requires_grad_(True)
        _, xyz, cov = G.pose_geometry(v, t(scene.faces), t(scene.lbs_weights), so3, sc, t(fr["cnl_gtfms"][b]),
                                      t(fr["dst_Rs"][b]), t(fr["dst_Ts"][b]))
        cov6 = G.pack_cov6(cov)
        st = Cam.raster_settings_from_KE(fr["K"][b], fr["E"][b], (W, H))
        feat = np.concatenate([app.detach().numpy().T, np.ones((scene.n_faces, 1), np.float32)], 1)
        f = R.forward(xyz.detach().numpy(), cov6.detach().numpy(), feat, np.ones(scene.n_faces, np.float32), st.viewmatrix,
                      st.projmatrix, st.tanfovx, st.tanfovy, np.zeros(4, np.float32), H, W)
        img = t(f["color"].transpose(1, 2, 0).copy())[None].requires_grad_(True)
        u = OL.unpack(img[..., :3], img[..., 3], t(fr["bgcolor"][b:b + 1]))
        l_rgb, l_mask = OL.l1_losses(u, img[..., 3], tgt, tgt_m)
        loss = l_rgb + 5.0 * l_mask + OL.lpips_loss(lp, u, tgt)
        loss.backward()
        g = R.backward(f, img.grad[0].permute(2, 0, 1).contiguous().numpy())
        ((xyz * t(g["means3D"])).sum() + (cov6 * t(g["cov6"])).sum()).backward()
        if gpu_device is not None and psnr is None:      # PSNR of the B200 render against the oracle render, same frame
            from gomavatar_b200.model import Model, default_model_cfg
            m = Model(default_model_cfg(img_size=(W, H)), scene.canonical_info()).to(gpu_device)   # same initial params
            with torch.no_grad():
                m.so3.copy_(t(pr["so3"])); m.scale.copy_(t(pr["scale"])); m.appearance_module.appearance.copy_(t(pr["appearance"]))
                d = {k: t(fr[k][b:b + 1]).to(gpu_device) for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts")}
                rgb, mask, _ = m(d["K"], d["E"], d["cnl_gtfms"], d["dst_Rs"], d["dst_Ts"])
            mse = float(((rgb[0].cpu().double() - img.detach()[0, ..., :3].double()) ** 2).mean())
            psnr = float("inf") if mse == 0 else -10.0 * np.log10(mse)
        return float(loss.detach())

    for _ in range(warmup):
        one_frame(0)
    t0 = time.perf_counter()
    for s in range(steps):
        for b in range(n_frames):
            one_frame(b)
    dt = time.perf_counter() - t0
    out = {"value": steps * n_frames / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{steps} step(s) x {n_frames} frame(s) of the same workload (512x512, {scene.n_faces} Gaussians, fwd+loss+bwd, "
                     f"batch 1 per frame like the reference); oracle port: torch-CPU ops + OpenMP C rasterizer ({R.num_threads()} threads)",
           "seconds": dt}
    if psnr is not None:
        out["psnr_b200_vs_oracle_db"] = psnr if np.isfinite(psnr) else 999.0
    return out


def run_reference(args):
    """`--impl reference`: the reference has no CPU-runnable implementation of this path (its rasterizer is a CUDA-only
    third-party package, pytorch3d is absent), so the oracle port stands in (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    K, W_ = args.steps, args.warmup
    t0 = time.perf_counter()
    res = cpu_path(args, n_frames=1, steps=K, warmup=min(W_, 1))
    ms_per_step = 1e3 * res["seconds"] / K
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": min(W_, 1), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (same generator and seeds as the B200 arm)",
            "config": {"workload": "same train step on the host cores, 1 frame per step (bounded sample of the B200 arm's batch)",
                       "img": args.img, "n_gaussians": args.faces, "frames_per_step_per_gpu": 1},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)
    return line


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: keep a private handle to it and point fd 1 at stderr, so that library chatter
    # written straight to fd 1 (NCCL prints its version banner there) cannot precede the line
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
