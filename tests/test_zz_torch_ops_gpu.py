"""torch.ops.mra.* on the CUDA kernels (bf16, tcgen05 path): the dispatcher-visible conv + norm pair against torch's own
modules in fp32, forward and backward (the CPU suite checks the same operators on the oracle ops, test_torch_ops_cpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from mra_gan_b200 import ops, torch_ops
from mra_gan_b200.ops import ACT_RELU
from oracle.functional import rel_l2

pytestmark = pytest.mark.gpu


def test_conv_norm_ops_on_the_gpu():
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((2, 64, 12, 12, 12), generator=gen).bfloat16().float().requires_grad_(True)
    w = (torch.randn((128, 64, 3, 3, 3), generator=gen) * 0.03).bfloat16().float().requires_grad_(True)
    b = torch.randn(128, generator=gen).requires_grad_(True)
    y = F.pad(F.relu(F.instance_norm(F.conv3d(x, w, b), eps=1e-5)), (1,) * 6, mode="replicate")
    gy = torch.randn(y.shape, generator=gen).bfloat16().float()
    y.backward(gy)

    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
    cf = lambda t: t.permute(0, 4, 1, 2, 3)
    xo = cl(x.detach()).bfloat16().cuda().requires_grad_(True)
    wo = torch_ops.pack_weight(w.detach()).bfloat16().cuda().requires_grad_(True)
    bo = b.detach().cuda().requires_grad_(True)
    h = torch.ops.mra.conv3d(xo, wo, bo, 3, 1, 0, False, 0, 0, 0.0)
    z, mean, rstd = torch.ops.mra.inorm_act_pad(h, None, None, 1, ACT_RELU, 0.0, -1, 1e-5)
    assert z.dtype == torch.bfloat16 and tuple(z.shape) == (2, 12, 12, 12, 128)
    assert rel_l2(cf(z.detach().float().cpu()), y.detach()) < 1e-2
    z.backward(cl(gy).bfloat16().cuda())
    assert ops.impl().tc_error() == 0
    # two bf16-stored layers back to back (the CPU oracle ops with bf16 storage measure 2.4e-2 here; layer-isolated
    # parity at 1e-2 is test_ops_gpu.py's job)
    assert rel_l2(cf(xo.grad.float().cpu()), x.grad) < 5e-2
    assert rel_l2(wo.grad.float().cpu(), torch_ops.pack_weight(w.grad)) < 5e-2
    # a conv bias in front of a train-mode InstanceNorm has a mathematically zero gradient
    assert float(bo.grad.abs().max()) < 1e-2 * float(gy.abs().sum()) ** 0.5
