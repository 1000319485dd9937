"""GPU: where does a GRAPH-REPLAYED step spend its time?  CUPTI kernel/memset records of one replayed
optimize_parameters() (two CUDA graphs): sum of device durations, number of nodes by kind, and the idle time
between consecutive records on the stream (the launch / dependency gaps the kernels cannot see).
Usage: python tools/profile_graph_step.py [batch] [netG]   ->  gpurun_out/graph_step_profile_*.txt"""
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
netG = sys.argv[2] if len(sys.argv) > 2 else "resnet_9blocks"
N3.set_default_compute_dtype(torch.bfloat16)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt(netG=netG))
    m.setup(bench.make_opt(netG=netG))
m.enable_cuda_graphs(warmup_steps=2)
A = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
B = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
for _ in range(5):
    m.set_input([A, B]); m.optimize_parameters()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.set_input([A, B]); m.optimize_parameters()
e1.record(); e1.synchronize()
wall = e0.elapsed_time(e1) / 5
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.set_input([A, B]); m.optimize_parameters()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = sum(e.time_range.end - e.time_range.start for e in evs)
span = evs[-1].time_range.end - evs[0].time_range.start
gaps, small = [], 0
by = {}
for a, b in zip(evs[:-1], evs[1:]):
    gaps.append(max(0.0, b.time_range.start - a.time_range.end))
for e in evs:
    d = e.time_range.end - e.time_range.start
    k = e.name.split("(")[0][-60:]
    c = by.setdefault(k, [0, 0.0])
    c[0] += 1; c[1] += d
    if d < 5.0:
        small += 1
out = []
out.append("graph-replayed step, batch %d: wall %.2f ms/step (events, 5 steps); profiled step: %d device records, sum of durations %.2f ms, "
           "first-to-last span %.2f ms, idle between records %.2f ms (%.1f us mean gap); records shorter than 5 us: %d"
           % (batch, wall, len(evs), tot / 1e3, span / 1e3, sum(gaps) / 1e3, sum(gaps) / max(1, len(gaps)), small))
gs = sorted(gaps)
out.append("gap percentiles us: p50 %.2f p90 %.2f p99 %.2f max %.1f" % (gs[len(gs) // 2], gs[int(len(gs) * .9)], gs[int(len(gs) * .99)], gs[-1]))
# gap attributed to the record that FOLLOWS it, by kind
gby = {}
for g, e in zip(gaps, evs[1:]):
    k = e.name.split("(")[0][-60:]
    c = gby.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += g
out.append("%-62s %6s %10s %10s" % ("kind", "count", "dur ms", "gap-before ms"))
for k, (n, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    out.append("%-62s %6d %10.3f %10.3f" % (k, n, d / 1e3, gby.get(k, [0, 0.0])[1] / 1e3))
txt = "\n".join(out)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/graph_step_profile_%sb%d.txt" % ("" if netG == "resnet_9blocks" else netG + "_", batch), "w").write(txt + "\n")
print(txt)
