"""GPU: host-side issue time of one optimize_parameters() (no sync) vs the step's GPU time."""
import contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N3.set_default_compute_dtype(torch.bfloat16)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt()); m.setup(bench.make_opt())
A = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
B = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
for _ in range(3):
    m.set_input([A, B]); m.optimize_parameters()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    m.set_input([A, B]); m.optimize_parameters()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host issue %.1f ms ; until GPU done %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
