"""GPU: wgrad kernel on a list of shapes: error flag, rel-L2 vs fp64 reference, time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ConvGeom
from oracle import ops_ref as R
I = ops.impl()
ref = R.RefImpl(torch.float64)
CASES = [
    (ConvGeom(256, 256, 3, 1, 0), (10, 10, 10), 2),
    (ConvGeom(128, 256, 4, 1, 1), (10, 10, 10), 1),
    (ConvGeom(128, 512, 4, 1, 1), (10, 10, 10), 1),
    (ConvGeom(256, 512, 4, 1, 1), (10, 10, 10), 1),
    (ConvGeom(128, 128, 3, 1, 1), (10, 10, 10), 1),
    (ConvGeom(256, 256, 3, 1, 0), (6, 6, 34), 1),
    (ConvGeom(256, 256, 3, 1, 0), (6, 10, 18), 2),
    (ConvGeom(128, 128, 3, 1, 0), (6, 10, 34), 1),
    (ConvGeom(64, 64, 3, 1, 0), (6, 10, 34), 1),
    (ConvGeom(128, 64, 4, 1, 1), (5, 6, 33), 1),
    (ConvGeom(256, 128, 3, 2, 1, True, 1), (6, 6, 6), 2),
    (ConvGeom(128, 64, 3, 2, 1, True, 1), (8, 8, 8), 1),
]
for g, dims, n in CASES:
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((n,) + dims + (g.cin,), generator=gen).to(torch.bfloat16)
    dy = torch.randn((n,) + g.out_dims(dims) + (g.cout,), generator=gen).to(torch.bfloat16)
    torch.cuda.synchronize(); t0 = time.time()
    dw, _ = I.conv_wgrad(x.cuda(), dy.cuda(), g)
    torch.cuda.synchronize(); dt = time.time() - t0
    err = I.tc_error()
    dw_ref, _ = ref.conv_wgrad(x.double(), dy.double(), g)
    e = float((dw.cpu().double() - dw_ref).norm() / dw_ref.norm())
    print("cin %3d cout %3d k%d s%d p%d %s dims %s n%d : flag %2d rel %.3e  %.1f ms  plan %s" % (
        g.cin, g.cout, g.k, g.stride, g.pad, "T" if g.transposed else "C", dims, n, err, e, dt * 1e3,
        ops.plan_describe(g, n, dims, 2)[15:18]), flush=True)
