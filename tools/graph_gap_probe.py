"""GPU: what does a CUDA-graph boundary cost?  Host time of each replay call, and device time of back-to-back
replays (host queued far ahead behind a spin kernel) against the sum of the graph's kernel durations."""
import contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mra_gan_b200 import networks3D as N3  # noqa: E402
from mra_gan_b200.models import create_model  # noqa: E402
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N3.set_default_compute_dtype(torch.bfloat16)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt()); m.setup(bench.make_opt())
m.enable_cuda_graphs(warmup_steps=2)
A = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
B = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
for _ in range(6):
    m.set_input([A, B]); m.optimize_parameters()
torch.cuda.synchronize()
G = m._graphs
def ev(): return torch.cuda.Event(enable_timing=True)
# host cost of the calls
torch.cuda.synchronize()
t0 = time.perf_counter(); G["gG"].replay(); t1 = time.perf_counter(); G["gD"].replay(); t2 = time.perf_counter()
torch.cuda.synchronize()
print("host: gG.replay() %.3f ms, gD.replay() %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
# device: back-to-back replays with the host far ahead
for trial in range(2):
    torch.cuda._sleep(int(4e8))          # ~0.2 s
    es = [ev() for _ in range(7)]
    es[0].record(); G["gG"].replay(); es[1].record(); G["gG"].replay(); es[2].record(); G["gD"].replay(); es[3].record()
    G["gD"].replay(); es[4].record(); G["gG"].replay(); es[5].record(); G["gD"].replay(); es[6].record()
    t_host = time.perf_counter()
    torch.cuda.synchronize()
    print("device ms: gG %.3f gG %.3f gD %.3f gD %.3f gG %.3f gD %.3f" % tuple(es[i].elapsed_time(es[i + 1]) for i in range(6)))
# the real step loop, host ahead
torch.cuda._sleep(int(4e8))
e0, e1 = ev(), ev()
e0.record()
for _ in range(5):
    m.set_input([A, B]); m.optimize_parameters()
e1.record(); torch.cuda.synchronize()
print("step loop (host ahead): %.3f ms/step" % (e0.elapsed_time(e1) / 5))
# host time of one whole optimize_parameters()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    m.set_input([A, B]); m.optimize_parameters()
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host time per optimize_parameters(): %.3f ms" % ((t1 - t0) / 5 * 1e3))
