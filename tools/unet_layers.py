"""GPU: per-layer fprop / dgrad / wgrad times of the 14 convolutions of the 7-down UNet generator (BASELINE config 4,
ngf = 64, 128^3).  Usage: python tools/unet_layers.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ConvGeom
I = ops.impl()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = [("d1 1->64", ConvGeom(1, 64, 4, 2, 1), 128), ("d2 64->128", ConvGeom(64, 128, 4, 2, 1), 64), ("d3 128->256", ConvGeom(128, 256, 4, 2, 1), 32),
     ("d4 256->512", ConvGeom(256, 512, 4, 2, 1), 16), ("d5 512->512", ConvGeom(512, 512, 4, 2, 1), 8), ("d6 512->512", ConvGeom(512, 512, 4, 2, 1), 4),
     ("d7 512->512", ConvGeom(512, 512, 4, 2, 1), 2),
     ("u7 512->512 T", ConvGeom(512, 512, 4, 2, 1, True, 0), 1), ("u6 1024->512 T", ConvGeom(1024, 512, 4, 2, 1, True, 0), 2),
     ("u5 1024->512 T", ConvGeom(1024, 512, 4, 2, 1, True, 0), 4), ("u4 1024->256 T", ConvGeom(1024, 256, 4, 2, 1, True, 0), 8),
     ("u3 512->128 T", ConvGeom(512, 128, 4, 2, 1, True, 0), 16), ("u2 256->64 T", ConvGeom(256, 64, 4, 2, 1, True, 0), 32),
     ("u1 128->1 T", ConvGeom(128, 1, 4, 2, 1, True, 0), 64)]
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
tot = [0.0, 0.0, 0.0]
print("batch %d; ms per launch (TFLOP/s) ; weight MB (bf16)" % N)
for name, g, d in L:
    dims = (d,) * 3
    od = g.out_dims(dims)
    x = torch.randn((N,) + dims + (g.cin,), device="cuda").to(torch.bfloat16)
    w = (torch.randn((g.taps, g.cout, g.cin), device="cuda") * 0.02).to(torch.bfloat16)
    dy = torch.randn((N,) + od + (g.cout,), device="cuda").to(torch.bfloat16)
    wT = I.pack_weight_t(w, torch.bfloat16)
    macs = N * g.cin * g.cout * g.taps * (od[0] ** 3 if not g.transposed else d ** 3)
    if g.transposed: macs = N * g.cin * g.cout * g.taps * d ** 3
    tf = timeit(lambda: I.conv_fprop(x, w, None, g, want_stats=g.cout > 1))
    td = timeit(lambda: I.conv_dgrad(dy, wT, g, dims))
    tw = timeit(lambda: I.conv_wgrad(x, dy, g))
    tot[0] += tf; tot[1] += td; tot[2] += tw
    print("%-16s fprop %.4f (%7.1f)  dgrad %.4f (%7.1f)  wgrad %.4f (%7.1f)   w %.1f MB  tc=%s" % (
        name, tf, 2 * macs / tf / 1e9, td, 2 * macs / td / 1e9, tw, 2 * macs / tw / 1e9, w.numel() * 2 / 1e6,
        [int(I.conv_uses_tensor_cores(g, N, dims, torch.bfloat16, k)) for k in range(3)]), flush=True)
print("sum: fprop %.3f dgrad %.3f wgrad %.3f ms ; per G pass (f+d+w) %.3f ms ; flag %d" % (tot[0], tot[1], tot[2], sum(tot), I.tc_error()))
