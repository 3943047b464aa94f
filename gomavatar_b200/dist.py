"""Frame-sharded data parallelism: one process per GPU, every parameter replicated, each rank renders its own frames,
ONE all-reduce of a flat fp32 gradient arena per step.  The reference has no multi-GPU path at all
(batch size is hard-wired to 1: configs/default.yaml:10, gaussian.py:24), so N-rank semantics are defined as the
1-rank step over the same N x B_local frames with mean-reduced losses.

NVLink 5 / NVSwitch makes all peers uniform, and the arena is <= 3.8 MB (vertices 3V | so3 3F | scale 3F | appearance
3F | MLPs), so the collective is latency-bound: a single NCCL call on the compute stream right after the last backward
kernel — no bucketing, no hierarchy.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class FlatArena:
    """Re-homes every trainable parameter of ``module`` (and its .grad) as views into two flat fp32 buffers."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.data = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.slices = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.data[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.data[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)        # autograd accumulates in place into the arena
            self.slices.append((off, k))
            off += k
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce_mean(self, group=None):
        """One collective per step.  Mean over ranks == gradient of the mean-over-all-frames loss."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            self.grad.div_(dist.get_world_size(group))

    def all_reduce_sum(self, group=None):
        """Sum over ranks only; returns the factor (1/world) the optimizer should apply (ArenaAdam.step(grad_scale=...))."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def broadcast_params(self, src=0, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.data, src=src, group=group)


class ArenaAdam:
    """torch.optim.Adam semantics (amsgrad off, weight decay 0) over a FlatArena in ONE kernel launch
    (csrc/adam.cu; replaces reference train.py:339 for the groups of models/model.py:305-324).  ``param_groups`` is the
    list ``Model.get_param_groups`` returns (dicts with 'params' and 'lr'; 'name' optional); every arena parameter must
    appear in exactly one group.  ``param_groups[i]['lr']`` may be edited between steps like torch's (the reference's
    exponential decay, train.py:166-175).  ``grad_scale`` multiplies the gradient first — pass 1/world_size and use
    ``FlatArena.all_reduce_sum`` to fold the mean of the data-parallel all-reduce into this launch."""

    def __init__(self, arena, param_groups, betas=(0.9, 0.999), eps=1e-8):
        from . import _lib
        self.arena, self.betas, self.eps = arena, betas, eps
        self.param_groups = [dict(g) for g in param_groups]
        owner = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                owner[id(p)] = gi
        self.segments = []                                   # (end offset, group index), merged when adjacent
        for p, (off, k) in zip(arena.params, arena.slices):
            if id(p) not in owner:
                raise ValueError("ArenaAdam: an arena parameter is in no param group")
            gi = owner[id(p)]
            if self.segments and self.segments[-1][1] == gi:
                self.segments[-1] = (off + k, gi)
            else:
                self.segments.append((off + k, gi))
        if len(self.segments) > _lib.ADAM_MAX_SEGMENTS:
            raise ValueError(f"ArenaAdam: more than {_lib.ADAM_MAX_SEGMENTS} learning-rate segments")
        self.exp_avg = torch.zeros_like(arena.data)
        self.exp_avg_sq = torch.zeros_like(arena.data)
        self.step_count = 0

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    def step(self, grad_scale=1.0):
        from . import _lib
        self.step_count += 1
        b1, b2 = self.betas
        a = _lib.GomAdamArgs(n=self.arena.numel, param=_lib.ptr(self.arena.data), grad=_lib.ptr(self.arena.grad),
                             exp_avg=_lib.ptr(self.exp_avg), exp_avg_sq=_lib.ptr(self.exp_avg_sq), beta1=b1, beta2=b2,
                             eps=self.eps, grad_scale=float(grad_scale), bias_correction1=1.0 - b1 ** self.step_count,
                             bias_correction2=1.0 - b2 ** self.step_count, n_segments=len(self.segments))
        for s, (end, gi) in enumerate(self.segments):
            a.seg_end[s] = end
            a.seg_lr[s] = float(self.param_groups[gi]["lr"])
        _lib.call("gom_adam_step", a)


def shard_frames(n_frames_global: int, rank: int, world: int):
    """Frames {rank, rank+world, ...} of the step's global batch (SURVEY.md §8e partitioning)."""
    return list(range(rank, n_frames_global, world))


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)      # binds the communicator to this rank's GPU (and barrier() to it)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world
