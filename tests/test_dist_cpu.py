"""CPU (gloo, world_size 2): the frame-sharded data-parallel plumbing of gomavatar_b200/dist.py — flat parameter /
gradient arena, frame sharding, ONE all-reduce per step — gives the same gradient and the same Adam step as a single
rank over all frames.  The hot-path kernels themselves need a GPU; here the 'model' is a small torch
module so that only the host-side logic is exercised."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gomavatar_b200.dist import FlatArena, init_from_env, shard_frames


class Toy(torch.nn.Module):
    """parameters shaped like the reference's SoA tensors ([3,V], [3,F]) + one frozen buffer-like parameter"""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.vertices = torch.nn.Parameter(torch.randn(3, 17, generator=g))
        self.so3 = torch.nn.Parameter(torch.randn(3, 29, generator=g))
        self.frozen = torch.nn.Parameter(torch.randn(5, generator=g), requires_grad=False)

    def forward(self, frame):                      # frame: [4] "pose" of one frame
        return (self.vertices * frame[0]).sin().sum() * frame[1] + (self.so3 ** 2).sum() * frame[2] + frame[3]


def _frames(n):
    return torch.from_numpy(np.random.default_rng(1).normal(size=(n, 4)).astype(np.float32))


def _step(model, arena, opt, frames, n_global):
    arena.zero_grad()
    for f in frames:
        (model(f) / len(frames)).backward()        # mean over the local frames
    arena.all_reduce_mean()                        # mean over ranks == mean over all frames (equal shard sizes)
    opt.step()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, _, w = init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                  # deliberately different initial replicas ...
    model = Toy()
    with torch.no_grad():
        if rank != 0:
            model.vertices.add_(1.0)
    arena = FlatArena(model)
    arena.broadcast_params(src=0)                  # ... made identical by the broadcast
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    frames = _frames(n_frames)
    mine = frames[shard_frames(n_frames, rank, world)]
    for _ in range(3):
        _step(model, arena, opt, mine, n_frames)
    out[rank] = (arena.data.clone(), arena.grad.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_arena_views_and_zero_grad():
    m = Toy()
    a = FlatArena(m)
    assert a.numel == 3 * 17 + 3 * 29                          # the frozen parameter is not in the arena
    assert m.vertices.data_ptr() == a.data.data_ptr() and m.vertices.grad.data_ptr() == a.grad.data_ptr()
    m(_frames(1)[0]).backward()
    assert float(a.grad.abs().sum()) > 0 and torch.equal(a.grad[: 3 * 17].view(3, 17), m.vertices.grad)
    a.zero_grad()
    assert float(m.so3.grad.abs().sum()) == 0
    assert shard_frames(8, 1, 4) == [1, 5] and sorted(sum((shard_frames(8, r, 4) for r in range(4)), [])) == list(range(8))


@pytest.mark.timeout(120)
def test_two_rank_step_equals_single_rank_step():
    n_frames, world = 8, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_frames, out), nprocs=world, join=True)
    # single-rank reference over all frames
    model = Toy()
    arena = FlatArena(model)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    for _ in range(3):
        _step(model, arena, opt, _frames(n_frames), n_frames)
    for rank in range(world):
        data, grad = out[rank]
        torch.testing.assert_close(grad, arena.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(data, arena.data, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out[0][0], out[1][0], rtol=0, atol=0)       # replicas stay bit-identical


def _subdivide_worker(rank, world, port, out):
    """a subdivision is a deterministic host event every rank executes at the same step — no communication.
    The real ``Model`` (host side only: no forward), a rank-dependent pseudo-gradient, one all-reduce before and one after."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    init_from_env("gloo")
    from gomavatar_b200 import synthetic as S
    from gomavatar_b200.model import Model, default_model_cfg
    model = Model(default_model_cfg((64, 64)), S.make_humanoid(2000, seed=0).canonical_info())
    log = []
    for phase in range(2):
        arena = FlatArena(model)                                   # re-created after the subdivision, like the optimizer
        arena.broadcast_params(src=0)
        g = torch.Generator().manual_seed(10 * phase + rank)
        arena.grad.copy_(torch.randn(arena.numel, generator=g))
        arena.all_reduce_mean()
        with torch.no_grad():
            arena.data.add_(arena.grad, alpha=-1e-3)               # a plain SGD step stands in for the Adam kernel
        log.append((arena.numel, arena.data.clone(), model.faces.clone(), model.lbs_weights.clone()))
        if phase == 0:
            model.subdivide()
    out[rank] = log
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_replicas_stay_identical_across_a_subdivision():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_subdivide_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (n0, d0, f0, w0), (n1, d1, f1, w1) = out[0]
    assert n1 == 3 * (1034 + 3000) + 9 * 8000 and n0 == 3 * 1034 + 9 * 2000 and f1.shape[0] == 4 * f0.shape[0]
    for phase in range(2):
        for a, b in zip(out[0][phase][1:], out[1][phase][1:]):
            assert torch.equal(a, b)                               # parameters, faces, LBS weights: bit-identical on both ranks
