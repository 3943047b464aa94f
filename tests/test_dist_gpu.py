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
