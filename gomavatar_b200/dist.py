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
    ``FlatArena.all_reduce_sum`` to fold the mean of the data-parallel all-reduce into this launch.

    Like torch.optim.Adam, every group keeps its OWN step counter and a group that received no gradient is skipped
    entirely — ``step(active=...)`` names the groups that were part of this step's graph (the reference's non-rigid /
    pose-refinement MLPs have ``.grad is None`` before their kick_in_iter; an arena gradient is zero instead, which must
    not age the moments or the bias correction).  ``state_dict`` / ``load_state_dict`` speak torch.optim.Adam's layout
    (per-parameter 'step' / 'exp_avg' / 'exp_avg_sq', ``param_groups`` with integer parameter ids in
    ``get_param_groups`` order), so the reference's ``train.py --resume`` (train.py:281) reads these checkpoints and this
    class reads the reference's.

    ``device_state=True`` keeps the counters in device memory and (with ``lr_decay=(rate, steps)``) evaluates the
    reference's exponential learning-rate decay on the device: the launch then has no step-dependent argument and can be
    captured in a CUDA graph together with the forward / backward (edits of ``param_groups[i]['lr']`` after the capture
    are not seen by the graph)."""

    def __init__(self, arena, param_groups, betas=(0.9, 0.999), eps=1e-8, device_state=False, lr_decay=None):
        from . import _lib
        self.arena, self.betas, self.eps = arena, betas, eps
        self.param_groups = [dict(g) for g in param_groups]
        self.lr_decay = lr_decay
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
        self.group_steps = [0] * len(self.param_groups)      # completed steps per group (host form)
        self.step_count = 0                                  # optimizer-wide
        self.dev_steps = (torch.zeros(_lib.ADAM_MAX_SEGMENTS + 1, dtype=torch.int64, device=arena.data.device)
                          if device_state else None)

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    def _active_flags(self, active):
        if active is None:
            return [True] * len(self.param_groups)
        want = set(active)
        return [(gi in want) or (g.get("name") in want) for gi, g in enumerate(self.param_groups)]

    def step(self, grad_scale=1.0, active=None):
        """``active``: group indices and / or names that received a gradient this step (None: all of them)."""
        from . import _lib
        flags = self._active_flags(active)
        b1, b2 = self.betas
        rate, steps = self.lr_decay if self.lr_decay else (1.0, 0.0)
        a = _lib.GomAdamArgs(n=self.arena.numel, param=_lib.ptr(self.arena.data), grad=_lib.ptr(self.arena.grad),
                             exp_avg=_lib.ptr(self.exp_avg), exp_avg_sq=_lib.ptr(self.exp_avg_sq), beta1=b1, beta2=b2,
                             eps=self.eps, grad_scale=float(grad_scale), lr_decay_rate=float(rate), lr_decay_steps=float(steps),
                             n_segments=len(self.segments), dev_steps=_lib.ptr(self.dev_steps), iter=self.step_count)
        for s, (end, gi) in enumerate(self.segments):
            a.seg_end[s] = end
            a.seg_lr[s] = float(self.param_groups[gi]["lr"])
            a.seg_active[s] = int(flags[gi])
            a.seg_step[s] = self.group_steps[gi] + 1
        _lib.call("gom_adam_step", a)
        for gi, f in enumerate(flags):
            if f:
                self.group_steps[gi] += 1
        self.step_count += 1

    # ------------------------------------------------------------------------------- torch.optim.Adam checkpoint layout
    def _sync_host_steps(self):
        if self.dev_steps is not None:                       # one read-back, only when a checkpoint is written
            d = self.dev_steps.cpu().tolist()
            for s, (_, gi) in enumerate(self.segments):
                self.group_steps[gi] = int(d[s])
            self.step_count = int(d[-1])

    def state_dict(self):
        self._sync_host_steps()
        where = {id(p): (off, k) for p, (off, k) in zip(self.arena.params, self.arena.slices)}
        state, groups, pid = {}, [], 0
        b1, b2 = self.betas
        for gi, g in enumerate(self.param_groups):
            ids = []
            for p in g["params"]:
                if id(p) in where and self.group_steps[gi] > 0:      # torch creates the state at a parameter's first step
                    off, k = where[id(p)]
                    state[pid] = {"step": torch.tensor(float(self.group_steps[gi])),
                                  "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                                  "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
                ids.append(pid)
                pid += 1
            meta = {k: v for k, v in g.items() if k != "params"}
            meta.update({"betas": (b1, b2), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                         "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": ids})
            groups.append(meta)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        if not sd or "param_groups" not in sd:
            raise KeyError("param_groups")                   # what torch.optim.Optimizer.load_state_dict raises on {}
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        where = {id(p): (off, k) for p, (off, k) in zip(self.arena.params, self.arena.slices)}
        state = sd.get("state", {})
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        for gi, (g, sg) in enumerate(zip(self.param_groups, sd["param_groups"])):
            if len(sg["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            g["lr"] = sg.get("lr", g["lr"])
            self.group_steps[gi] = 0
            for p, pid in zip(g["params"], sg["params"]):
                st = state.get(pid, state.get(str(pid)))
                if st is None or id(p) not in where:
                    continue
                off, k = where[id(p)]
                self.exp_avg[off:off + k].copy_(torch.as_tensor(st["exp_avg"]).reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(torch.as_tensor(st["exp_avg_sq"]).reshape(-1))
                self.group_steps[gi] = max(self.group_steps[gi], int(float(st["step"])))
        self.step_count = max(self.group_steps) if self.group_steps else 0
        if self.dev_steps is not None:
            d = torch.zeros_like(self.dev_steps, device="cpu")
            for s, (_, gi) in enumerate(self.segments):
                d[s] = self.group_steps[gi]
            d[-1] = self.step_count
            self.dev_steps.copy_(d)


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
