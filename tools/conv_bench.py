"""GPU: per-layer timing of every convolution of the BASELINE model (resnet_9blocks G + 3-layer D,
ngf=ndf=64) at 128^3: fprop / dgrad / wgrad ms and TFLOP/s, plus each layer's share of a training
step (weighted by how often it runs per optimize_parameters()).  Writes gpurun_out/conv_bench.txt.
Usage: python tools/conv_bench.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ConvGeom  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
# name, geom, input dims, (fprop, dgrad, wgrad) launches per training step
G = 6  # generator passes per step
LAYERS = [
    ("G.c1 1->64 k7", ConvGeom(1, 64, 7, 1, 0), (134,) * 3, (G, 2, G)),
    ("G.d1 64->128 k3s2", ConvGeom(64, 128, 3, 2, 1), (128,) * 3, (G, G, G)),
    ("G.d2 128->256 k3s2", ConvGeom(128, 256, 3, 2, 1), (64,) * 3, (G, G, G)),
    ("G.rb 256->256 k3", ConvGeom(256, 256, 3, 1, 0), (34,) * 3, (18 * G, 18 * G, 18 * G)),
    ("G.u1 256->128 T k3s2", ConvGeom(256, 128, 3, 2, 1, True, 1), (32,) * 3, (G, G, G)),
    ("G.u2 128->64 T k3s2", ConvGeom(128, 64, 3, 2, 1, True, 1), (64,) * 3, (G, G, G)),
    ("G.c4 64->1 k7", ConvGeom(64, 1, 7, 1, 0), (134,) * 3, (G, G, G)),
    ("D.1 1->64 k4s2", ConvGeom(1, 64, 4, 2, 1), (128,) * 3, (6, 2, 4)),
    ("D.2 64->128 k4s2", ConvGeom(64, 128, 4, 2, 1), (64,) * 3, (6, 6, 4)),
    ("D.3 128->256 k4s2", ConvGeom(128, 256, 4, 2, 1), (32,) * 3, (6, 6, 4)),
    ("D.4 256->512 k4s1", ConvGeom(256, 512, 4, 1, 1), (16,) * 3, (6, 6, 4)),
    ("D.5 512->1 k4s1", ConvGeom(512, 1, 4, 1, 1), (15,) * 3, (6, 6, 4)),
]


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    I = ops.impl()
    rows, total = [], 0.0
    for name, g, dims, mult in LAYERS:
        x = torch.randn((N,) + dims + (g.cin,), device="cuda").to(torch.bfloat16)
        w = (torch.randn((g.taps, g.cout, g.cin), device="cuda") * 0.02).to(torch.bfloat16)
        dy = torch.randn((N,) + g.out_dims(dims) + (g.cout,), device="cuda").to(torch.bfloat16)
        wT = I.pack_weight_t(w, torch.bfloat16)
        odims = g.out_dims(dims)
        macs = N * g.cin * g.cout * g.taps * (odims[0] * odims[1] * odims[2] if not g.transposed else dims[0] * dims[1] * dims[2])
        flops = 2.0 * macs
        t_f = timeit(lambda: I.conv_fprop(x, w, None, g, want_stats=g.cout > 1))
        t_d = timeit(lambda: I.conv_dgrad(dy, wT, g, dims))
        t_w = timeit(lambda: I.conv_wgrad(x, dy, g))
        step_ms = t_f * mult[0] + t_d * mult[1] + t_w * mult[2]
        total += step_ms
        rows.append((name, t_f, t_d, t_w, flops, step_ms))
        del x, w, dy, wT
        torch.cuda.empty_cache()
    out = ["batch %d, 128^3 patch; ms per launch (TFLOP/s); step share = launches/step x ms" % N,
           "%-22s %18s %18s %18s %10s %7s" % ("layer", "fprop", "dgrad", "wgrad", "step ms", "share")]
    for name, t_f, t_d, t_w, fl, sm in rows:
        f = lambda t: "%7.3f (%6.1f)" % (t, fl / t / 1e9)
        out.append("%-22s %18s %18s %18s %10.2f %6.1f%%" % (name, f(t_f), f(t_d), f(t_w), sm, 100 * sm / total))
    out.append("total conv time per step: %.1f ms ; tc error flag %d" % (total, I.tc_error()))
    txt = "\n".join(out)
    print(txt)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/conv_bench.txt", "w").write(txt + "\n")


if __name__ == "__main__":
    main()
