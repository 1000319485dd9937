"""GPU: launch the dominant kernels (Conv3d 256->256 k3 on 34^3 = G.rb: gather_halo_kernel fprop / dgrad,
wgrad_tc_kernel<pair>, the norm kernels around them) a few times -- the target of `ncu --set full` captures.

    python tools/one_kernel.py [batch] [iterations]

Per iteration, in this order, 8 kernels matching regex:gather_halo|wgrad_tc|inorm_ :
  0 gather_halo (fprop + statistics)   1 inorm_fwd_stream   2 inorm_bwd_stats_stream   3 inorm_bwd_stream
  4 gather_halo (dgrad)                5 wgrad_tc           6 gather_halo<.., aux> (dgrad + norm-backward statistics)
  7 inorm_bwd_stream (apply only)
so `ncu --set full -k regex:"gather_halo|wgrad_tc|inorm_" --launch-skip 8*(iterations-1) --launch-count 8` captures
the last iteration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ACT_RELU, ConvGeom  # noqa: E402

I = ops.impl()
g = ConvGeom(256, 256, 3, 1, 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = torch.randn((N, 34, 34, 34, 256), device="cuda").to(torch.bfloat16)
w = (torch.randn((27, 256, 256), device="cuda") * 0.02).to(torch.bfloat16)
wT = I.pack_weight_t(w, torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(ITERS):
    flush.zero_()
    y, st = I.conv_fprop(x, w, None, g, want_stats=True)
    z, mean, rstd = I.inorm_fwd(y, st, None, 1, ACT_RELU)
    dy, _ = I.inorm_bwd(z, y, mean, rstd, 1, ACT_RELU)
    flush.zero_()
    dx = I.conv_dgrad(dy, wT, g, (34, 34, 34))
    flush.zero_()
    dw, _ = I.conv_wgrad(x, dy, g)
    flush.zero_()
    dx2, sums = I.conv_dgrad_nstats(dy, wT, g, (34, 34, 34), z, ACT_RELU, 0.0)
    dz, _ = I.inorm_bwd_apply(dx2, y, mean, rstd, sums, 1, ACT_RELU)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), I.tc_error())
