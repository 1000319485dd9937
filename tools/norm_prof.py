"""GPU: one eager pass of the norm kernels over the two shapes that matter (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ACT_RELU, ACT_NONE
I = ops.impl()
for S, Cc, pad, res in ((128, 64, 0, False), (32, 256, 1, False), (32, 256, 1, True), (128, 64, 3, False)):
    x = torch.randn((2, S, S, S, Cc), device="cuda").to(torch.bfloat16)
    stats = I.inorm_stats(x)
    r = torch.randn((2, S + 2, S + 2, S + 2, Cc), device="cuda").to(torch.bfloat16) if res else None
    act = ACT_NONE if res else ACT_RELU
    for _ in range(2):
        y, mean, rstd = I.inorm_fwd(x, stats, r, pad, act, 0.0, 1 if res else -1)
        gy = torch.randn_like(y) if _ == 0 else gy
        dx, dres = I.inorm_bwd(gy, x, mean, rstd, pad, act, 0.0, 1 if res else -1)
    torch.cuda.synchronize()
