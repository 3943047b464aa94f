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

    def broadcast_params(self, src=0, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.data, src=src, group=group)


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
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world
