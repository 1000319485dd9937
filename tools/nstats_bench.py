"""GPU: what the norm-backward statistics cost inside the dgrad epilogue (mra_conv3d_dgrad_nstats) against the
statistics kernel they replace, per linked layer of the BASELINE model at 128^3.  Usage: python tools/nstats_bench.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ACT_LRELU, ACT_RELU, ConvGeom  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
# consumer conv, its input dims (= stored output of the norm, halo included), halo of the norm, act, count per step
LAYERS = [
    ("G.d1 dgrad <- norm(c1) 64ch 128^3", ConvGeom(64, 128, 3, 2, 1), (128,) * 3, 0, ACT_RELU, 6),
    ("G.d2 dgrad <- norm(d1) 128ch 64^3", ConvGeom(128, 256, 3, 2, 1), (64,) * 3, 0, ACT_RELU, 6),
    ("G.rb conv2 dgrad <- norm1 256ch 34^3", ConvGeom(256, 256, 3, 1, 0), (34,) * 3, 1, ACT_RELU, 54),
    ("G.u2 dgrad <- norm(u1) 128ch 64^3", ConvGeom(128, 64, 3, 2, 1, True, 1), (64,) * 3, 0, ACT_RELU, 6),
    ("G.c4 dgrad <- norm(u2) 64ch 134^3", ConvGeom(64, 1, 7, 1, 0), (134,) * 3, 3, ACT_RELU, 6),
    ("D.3 dgrad <- norm(D.2) 128ch 32^3", ConvGeom(128, 256, 4, 2, 1), (32,) * 3, 0, ACT_LRELU, 6),
    ("D.4 dgrad <- norm(D.3) 256ch 16^3", ConvGeom(256, 512, 4, 1, 1), (16,) * 3, 0, ACT_LRELU, 6),
    ("D.5 dgrad <- norm(D.4) 512ch 15^3", ConvGeom(512, 1, 4, 1, 1), (15,) * 3, 0, ACT_LRELU, 6),
]


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


I = ops.impl()
tot = [0.0, 0.0]
print("batch %d; us per call" % N)
print("%-40s %9s %9s %9s %9s %9s   %s" % ("layer", "dgrad", "+nstats", "stats k.", "apply k.", "2-pass", "gain/step us"))
for name, g, dims, pad, act, cnt in LAYERS:
    slope = 0.2 if act == ACT_LRELU else 0.0
    inner = tuple(d - 2 * pad for d in dims)
    x = torch.randn((N,) + inner + (g.cin,), device="cuda").to(torch.bfloat16)
    st = I.inorm_stats(x)
    y, mean, rstd = I.inorm_fwd(x, st, None, pad, act, slope, -1)
    dy = torch.randn((N,) + g.out_dims(dims) + (g.cout,), device="cuda").to(torch.bfloat16)
    w = (torch.randn((g.taps, g.cout, g.cin), device="cuda") * 0.02).to(torch.bfloat16)
    wT = I.pack_weight_t(w, torch.bfloat16)
    low, ws = I.conv_shared_workspace(g, N, dims, torch.bfloat16, "cuda")
    kw = dict(ws=ws) if low else {}
    ok = I.conv_dgrad_nstats_supported(g, N, dims, torch.bfloat16)
    t_d = timeit(lambda: I.conv_dgrad(dy, wT, g, dims, **kw))
    t_n = timeit(lambda: I.conv_dgrad_nstats(dy, wT, g, dims, y, act, slope, **kw)) if ok else float("nan")
    gy = I.conv_dgrad(dy, wT, g, dims, **kw)
    sums = I.inorm_bwd_stats(gy, x, mean, rstd, pad, act, slope, -1)
    t_s = timeit(lambda: I.inorm_bwd_stats(gy, x, mean, rstd, pad, act, slope, -1))
    t_a = timeit(lambda: I.inorm_bwd_apply(gy, x, mean, rstd, sums, pad, act, slope, -1))
    t_2 = timeit(lambda: I.inorm_bwd(gy, x, mean, rstd, pad, act, slope, -1))
    gain = (t_s - (t_n - t_d)) * cnt
    tot[0] += gain
    print("%-40s %9.1f %9.1f %9.1f %9.1f %9.1f   %+9.0f" % (name, t_d, t_n, t_s, t_a, t_2, gain), flush=True)
    del x, y, dy, gy
    torch.cuda.empty_cache()
print("sum of gains per step: %.0f us ; tc error %d" % (tot[0], I.tc_error()))
