"""GPU, world_size 2 over NCCL: N ranks compute the 1-rank gradient (SURVEY.md §4 "Multi-GPU tests", §8e).  The real Model,
kernels, losses and arena Adam run on two GPUs of the box (tests/host_harness/dist_grad_check.py under torch.distributed.run);
skipped where fewer than two GPUs are visible.  Tolerance 1e-4 of each parameter's largest gradient entry: the two sides add
the same per-frame contributions in a different order (atomics inside a rank, the all-reduce across ranks)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gradient_equals_one_rank_gradient():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "host_harness", "dist_grad_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["world"] == 2 and out["replicas_identical_after_adam"]
    assert max(out["grad_rel_err_vs_1rank"].values()) < 1e-4, out


def test_split_batch_gradient_equals_whole_batch_gradient_on_one_gpu():
    """The arithmetic of the 2-rank test without NCCL, on one GPU (tools/batch_split_check.py): the mean of the gradients over
    frames {0,2} and {1,3} equals the gradient over {0,1,2,3} — per-image results of every kernel must not depend on which other
    frames share a launch.  With GOM_CONV_KSPLIT=0 (no K-split in the convolutions: every unsplit tile shape, single CTA or
    CTA pair, is bit-identical) the agreement is at fp32 summation-order level; with K-splits allowed the small deep VGG layers
    pick different splits for 4 and 8 images and ReLU decisions near zero move the gradient by ~1e-3 — bounded here at 2e-2."""
    tool = os.path.join(ROOT, "tools", "batch_split_check.py")
    out = {}
    for flag in ("--no-ksplit", None):
        cmd = [sys.executable, tool] + ([flag] if flag else [])
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
        out[flag] = json.loads(lines[-1])
    assert max(out["--no-ksplit"]["rel"].values()) < 1e-5, out["--no-ksplit"]
    assert max(out[None]["rel"].values()) < 2e-2, out[None]
