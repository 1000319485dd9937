"""GPU: the head layer's fprop (Conv3d 64 -> 1 k7 on a replication-padded 134^3 tensor, batch 2) three times -- the target of an
`ncu --set full -k regex:shift_sum` capture of shift_sum_kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ACT_TANH, ConvGeom  # noqa: E402

I = ops.impl()
g = ConvGeom(64, 1, 7, 1, 0)
x = torch.randn((2, 134, 134, 134, 64), device="cuda").to(torch.bfloat16)
w = (torch.randn((343, 1, 64), device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(1, device="cuda")
for _ in range(3):
    y, _ = I.conv_fprop(x, w, b, g, act=ACT_TANH)
torch.cuda.synchronize()
print("ok", tuple(y.shape), I.tc_error())
