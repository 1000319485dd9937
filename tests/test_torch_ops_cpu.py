"""torch.ops.mra.* (mra_gan_b200/torch_ops.py): schemas, fake-tensor functions and autograd formulas of the
dispatcher-visible operators, against the torch.nn.functional calls the reference's modules make
(models/networks3D.py:186-213,241-257).  CPU: the oracle ops are installed as the implementation."""
import pytest
import torch
import torch.nn.functional as F

from mra_gan_b200 import ops, torch_ops
from mra_gan_b200.ops import ACT_LRELU, ACT_NONE, ACT_RELU
from oracle import ops_ref as R
from oracle.functional import rel_l2


@pytest.fixture(autouse=True)
def oracle_impl():
    prev = ops.set_impl(R.RefImpl(torch.float64))
    yield
    ops.set_impl(prev)


CL = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()      # NCDHW -> channels-last
CF = lambda t: t.permute(0, 4, 1, 2, 3)

CASES = [  # cin, cout, k, stride, pad, transposed, output_padding, act, dims
    (3, 5, 3, 1, 0, False, 0, ACT_NONE, (7, 6, 8)),
    (4, 6, 3, 2, 1, False, 0, ACT_LRELU, (8, 9, 8)),
    (4, 2, 4, 2, 1, False, 0, ACT_NONE, (8, 8, 10)),
    (4, 3, 3, 2, 1, True, 1, ACT_RELU, (4, 5, 4)),
    (3, 2, 4, 2, 1, True, 0, ACT_NONE, (5, 4, 4)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "c%d-%d_k%d_s%d_p%d_%s" % (c[0], c[1], c[2], c[3], c[4], "T" if c[5] else "C"))
def test_conv3d_op_matches_torch_and_differentiates(case):
    cin, cout, k, s, p, tr, op, act, dims = case
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((2, cin) + dims, generator=gen, dtype=torch.float64, requires_grad=True)
    wshape = (cin, cout, k, k, k) if tr else (cout, cin, k, k, k)
    w = (torch.randn(wshape, generator=gen, dtype=torch.float64) * 0.2).requires_grad_(True)
    b = torch.randn(cout, generator=gen, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose3d(x, w, b, stride=s, padding=p, output_padding=op) if tr else F.conv3d(x, w, b, stride=s, padding=p)
    y = F.relu(y) if act == ACT_RELU else (F.leaky_relu(y, 0.2) if act == ACT_LRELU else y)
    gy = torch.randn(y.shape, generator=gen, dtype=torch.float64)
    y.backward(gy)

    xo = CL(x.detach()).requires_grad_(True)
    wo = torch_ops.pack_weight(w.detach(), tr).requires_grad_(True)
    bo = b.detach().clone().requires_grad_(True)
    yo = torch.ops.mra.conv3d(xo, wo, bo, k, s, p, tr, op, act, 0.2)
    assert rel_l2(CF(yo.detach()), y.detach()) < 1e-12
    yo.backward(CL(gy))
    assert rel_l2(CF(xo.grad), x.grad) < 1e-12
    assert rel_l2(wo.grad, torch_ops.pack_weight(w.grad, tr)) < 1e-6          # wgrad is an fp32 result
    assert rel_l2(bo.grad, b.grad) < 1e-6
    # without a bias, and gradients only where asked
    y2 = torch.ops.mra.conv3d(xo.detach(), wo, None, k, s, p, tr, op, ACT_NONE, 0.0)
    (g2,) = torch.autograd.grad(y2.sum(), wo)
    assert g2.shape == wo.shape


def test_conv3d_stats_feeds_the_norm_op():
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((2, 4, 8, 7, 6), generator=gen, dtype=torch.float64)
    w = torch.randn((6, 4, 3, 3, 3), generator=gen, dtype=torch.float64) * 0.2
    b = torch.randn(6, generator=gen, dtype=torch.float64)
    y, st = torch.ops.mra.conv3d_stats(CL(x), torch_ops.pack_weight(w), b, 3, 1, 0, False, 0)
    want = F.conv3d(x, w, b)
    assert rel_l2(CF(y), want) < 1e-12
    assert rel_l2(st[..., 0], want.sum((2, 3, 4))) < 1e-12 and rel_l2(st[..., 1], (want * want).sum((2, 3, 4))) < 1e-12
    z, mean, rstd = torch.ops.mra.inorm_act_pad(y, st, None, 1, ACT_RELU, 0.0, -1, 1e-5)
    zr = F.pad(F.relu(F.instance_norm(want, eps=1e-5)), (1,) * 6, mode="replicate")
    assert rel_l2(CF(z), zr) < 1e-6                                             # mean / rstd travel as fp32
    assert rel_l2(mean, want.mean((2, 3, 4))) < 1e-6


@pytest.mark.parametrize("act,pad,with_res", [(ACT_RELU, 1, False), (ACT_NONE, 1, True), (ACT_LRELU, 0, False)])
def test_inorm_act_pad_op_differentiates(act, pad, with_res):
    gen = torch.Generator().manual_seed(2)
    x = (torch.randn((2, 5, 6, 5, 7), generator=gen, dtype=torch.float64) * 1.5 + 0.3).requires_grad_(True)
    res = torch.randn((2, 5, 8, 7, 9), generator=gen, dtype=torch.float64, requires_grad=True) if with_res else None
    y = F.instance_norm(x, eps=1e-5)
    y = F.relu(y) if act == ACT_RELU else (F.leaky_relu(y, 0.2) if act == ACT_LRELU else y)
    if with_res:
        y = y + res[:, :, 1:-1, 1:-1, 1:-1]
    if pad:
        y = F.pad(y, (pad,) * 6, mode="replicate")
    gy = torch.randn(y.shape, generator=gen, dtype=torch.float64)
    y.backward(gy)
    xo = CL(x.detach()).requires_grad_(True)
    ro = CL(res.detach()).requires_grad_(True) if with_res else None
    yo, mean, rstd = torch.ops.mra.inorm_act_pad(xo, None, ro, pad, act, 0.2, 1 if with_res else -1, 1e-5)
    assert not mean.requires_grad and not rstd.requires_grad
    assert rel_l2(CF(yo.detach()), y.detach()) < 1e-6
    yo.backward(CL(gy))
    assert rel_l2(CF(xo.grad), x.grad) < 1e-5
    if with_res:
        assert rel_l2(CF(ro.grad), res.grad) < 1e-12


def test_opcheck_schemas_fake_tensors_and_autograd_registration():
    from torch.library import opcheck
    gen = torch.Generator().manual_seed(3)
    x = torch.randn((1, 5, 5, 5, 3), generator=gen, dtype=torch.float64, requires_grad=True)
    w = torch.randn((27, 4, 3), generator=gen, dtype=torch.float64, requires_grad=True)
    b = torch.randn(4, generator=gen, dtype=torch.float64, requires_grad=True)
    utils = ("test_schema", "test_faketensor", "test_autograd_registration")
    opcheck(torch.ops.mra.conv3d.default, (x, w, b, 3, 1, 0, False, 0, ACT_RELU, 0.0), test_utils=utils)
    opcheck(torch.ops.mra.conv3d_stats.default, (x.detach(), w.detach(), None, 3, 1, 0, False, 0), test_utils=utils)
    dy = torch.randn((1, 3, 3, 3, 4), generator=gen, dtype=torch.float64)
    opcheck(torch.ops.mra.conv3d_dgrad.default, (dy, w.detach().transpose(1, 2).contiguous(), [5, 5, 5], 3, 1, 0, False, 0), test_utils=utils)
    opcheck(torch.ops.mra.conv3d_wgrad.default, (x.detach(), dy, 3, 1, 0, False, 0), test_utils=utils)
    opcheck(torch.ops.mra.inorm_act_pad.default, (x, None, None, 1, ACT_RELU, 0.0, -1, 1e-5), test_utils=utils)
    mean, rstd = torch.zeros(1, 3), torch.ones(1, 3)
    gy = torch.randn((1, 7, 7, 7, 3), generator=gen, dtype=torch.float64)
    opcheck(torch.ops.mra.inorm_act_pad_bwd.default, (gy, x.detach(), mean, rstd, 1, ACT_RELU, 0.0, -1), test_utils=utils)


def test_ops_refuse_to_run_without_the_cuda_library():
    """no CPU fallback: with the product implementation selected, a CPU call fails loudly."""
    ops.set_impl(None)
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    x = torch.zeros((1, 4, 4, 4, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.mra.conv3d(x, torch.zeros((27, 2, 2)), None, 3, 1, 0, False, 0, ACT_NONE, 0.0)
