import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
N3.set_default_compute_dtype(torch.bfloat16)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt(netG="unet_128")); m.setup(bench.make_opt(netG="unet_128"))
A = torch.rand(2, 1, 128, 128, 128, device="cuda") * 2 - 1
B = torch.rand(2, 1, 128, 128, 128, device="cuda") * 2 - 1
for _ in range(2):
    m.set_input([A, B]); m.optimize_parameters()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.set_input([A, B]); m.optimize_parameters(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
tot = sum(e.device_time_total for e in prof.key_averages())
print("total GPU time %.1f ms" % (tot / 1e3))
for e in rows:
    print("%-80s %9.2f ms x%d" % (e.key[:80], e.device_time_total / 1e3, e.count))
