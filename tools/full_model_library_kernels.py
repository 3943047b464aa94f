"""Which torch / library kernels are left in the full reference-shaped step (bench.py --full-model), by CPU op that launched
them: torch.profiler over 2 eager steps."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

a = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="b200", frames_per_step=8, faces=30000, img=512, pool_steps=2,
                          lpips_precision="tf32", lpips_torch=False, lpips_epilogue="cudnn", lpips_conv="tcgen05", lpips_streams=1, no_extras=True,
                          cuda_graph=False, full_model=True, no_cpu_baseline=True, cpu_frames=1)
dev = torch.device("cuda:0")
tr = bench.Trainer(a, 0, 1, dev)
for i in range(2):
    tr.step_device(i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=False, record_shapes=True) as prof:
    for i in range(2):
        tr.step_device(i)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_device_time_total", row_limit=45, max_name_column_width=60, max_shapes_column_width=70))
