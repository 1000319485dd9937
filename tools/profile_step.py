"""GPU: per-kernel time table of one CycleGAN step (torch.profiler / CUPTI), written to
gpurun_out/step_profile.txt.  Usage: python tools/profile_step.py [batch] [patch]"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mra_gan_b200 import networks3D as N3  # noqa: E402
from mra_gan_b200.models import create_model  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
N3.set_default_compute_dtype(torch.bfloat16)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt())
    m.setup(bench.make_opt())
A = torch.rand(batch, 1, patch, patch, patch, device="cuda") * 2 - 1
B = torch.rand(batch, 1, patch, patch, patch, device="cuda") * 2 - 1
for _ in range(2):
    m.set_input([A, B]); m.optimize_parameters()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    m.set_input([A, B]); m.optimize_parameters()
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/step_profile.txt", "w").write(txt)
print(txt[-6000:])
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
