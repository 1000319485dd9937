"""GPU: launch the dominant kernel (Conv3d 256->256 k3 on 34^3, batch 2: gather_tc_kernel) and the
matching wgrad / norm kernels a few times -- the target of `ncu --set full` captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ACT_RELU, ConvGeom  # noqa: E402

I = ops.impl()
g = ConvGeom(256, 256, 3, 1, 0)
N = 2
x = torch.randn((N, 34, 34, 34, 256), device="cuda").to(torch.bfloat16)
w = (torch.randn((27, 256, 256), device="cuda") * 0.02).to(torch.bfloat16)
wT = I.pack_weight_t(w, torch.bfloat16)
for _ in range(6):
    y, st = I.conv_fprop(x, w, None, g, want_stats=True)
    z, mean, rstd = I.inorm_fwd(y, st, None, 1, ACT_RELU)
    dy, _ = I.inorm_bwd(z, y, mean, rstd, 1, ACT_RELU)
    dx = I.conv_dgrad(dy, wT, g, (34, 34, 34))
    dw, _ = I.conv_wgrad(x, dy, g)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), I.tc_error())
