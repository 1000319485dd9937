"""GPU: time one conv layer's fprop/dgrad/wgrad (bf16).  Usage: python tools/layer_bench.py cin cout k stride pad transposed(0/1) outpad dim [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ConvGeom
a = [int(v) for v in sys.argv[1:]]
g = ConvGeom(a[0], a[1], a[2], a[3], a[4], bool(a[5]), a[6]); dims = (a[7],) * 3; N = a[8] if len(a) > 8 else 2
I = ops.impl()
x = torch.randn((N,) + dims + (g.cin,), device="cuda").to(torch.bfloat16)
w = (torch.randn((g.taps, g.cout, g.cin), device="cuda") * 0.02).to(torch.bfloat16)
dy = torch.randn((N,) + g.out_dims(dims) + (g.cout,), device="cuda").to(torch.bfloat16)
wT = I.pack_weight_t(w, torch.bfloat16)
od = g.out_dims(dims)
macs = N * g.cin * g.cout * g.taps * (od[0] * od[1] * od[2] if not g.transposed else dims[0] * dims[1] * dims[2])
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
for name, fn in (("fprop", lambda: I.conv_fprop(x, w, None, g, want_stats=(g.cout > 1 and os.environ.get("STATS", "1") == "1"))), ("dgrad", lambda: I.conv_dgrad(dy, wT, g, dims)),
                 ("wgrad", lambda: I.conv_wgrad(x, dy, g))):
    if os.environ.get("ONLY") and os.environ["ONLY"] != name: continue
    I.debug_counters()
    t = timeit(fn)
    c = I.debug_counters()
    if c[5]:
        print("   per CTA: producer wait %.0f / total %.0f clk ; mma wait %.0f / total %.0f clk ; epilogue %.0f clk ; CTAs %d" % (
            c[0] / c[5], c[1] / c[5], c[2] / c[5], c[3] / c[5], c[4] / c[5], c[5]))
        print("   (col kernel: tiles per CTA %.0f)" % (c[1] / c[5]))
        print("   mma wait for a free accumulator %.0f clk ; mma wait for a halo plane %.0f clk" % (c[6] / c[5], c[7] / c[5]))
    print("%s %.4f ms  %.1f TFLOP/s  (flag %d)" % (name, t, 2 * macs / t / 1e9, I.tc_error()), flush=True)
if os.environ.get("KERNELS"):
    from torch.profiler import ProfilerActivity, profile
    for name, fn in (("fprop", lambda: I.conv_fprop(x, w, None, g, want_stats=(g.cout > 1 and os.environ.get("STATS", "1") == "1"))), ("dgrad", lambda: I.conv_dgrad(dy, wT, g, dims)),
                     ("wgrad", lambda: I.conv_wgrad(x, dy, g))):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        print("--", name)
        for e in prof.key_averages():
            if e.device_time_total > 0:
                print("   %-70s %8.1f us x%d" % (e.key[:70], e.device_time_total / max(1, e.count), e.count))
