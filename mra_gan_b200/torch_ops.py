"""``torch.library`` registration of the C-ABI ops: ``torch.ops.mra.*`` (last row).

The package's own networks call ``ops.impl()`` directly from their ``autograd.Function``s (functional.py) -- one Python
frame less per launch.  This module is the same kernels as dispatcher-visible operators, for a maintainer who wants to
replace individual ``nn.Conv3d`` / ``nn.InstanceNorm3d`` call sites of the reference (models/networks3D.py:186-213,
241-257, 392-417) without adopting the fused networks: schemas, fake-tensor (shape) functions and autograd formulas are
registered, so the ops compose with ``torch.autograd``, ``torch.library.opcheck`` and tracing.

Conventions are the C ABI's (include/mra_gan_b200.h): activations channels-last ``(N, D, H, W, C)`` in bf16 or fp32,
weights packed ``[taps][Cout][Cin]`` in the activations' dtype (``pack_weight`` below converts the
reference's ``(Cout, Cin, k, k, k)`` / ``(Cin, Cout, k, k, k)`` parameters).  Every implementation function launches
through ``ops.impl()``: the CUDA library, or nothing -- there is no CPU fallback (the CPU test-suite installs the
oracle ops with ``ops.set_impl``).
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import ops
from .ops import ACT_NONE, ConvGeom


def pack_weight(w_ref: Tensor, transposed: bool = False) -> Tensor:
    """reference parameter layout -> [taps][Cout][Cin] (a copy; same dtype)."""
    k = w_ref.shape[2]
    p = w_ref.permute(*((2, 3, 4, 1, 0) if transposed else (2, 3, 4, 0, 1))).contiguous()
    return p.view(k ** 3, p.shape[3], p.shape[4])


def _geom(w: Tensor, k: int, stride: int, pad: int, transposed: bool, output_padding: int, wt: bool = False) -> ConvGeom:
    taps, a, b = w.shape
    if taps != k ** 3:
        raise ValueError("packed weight has %d taps, kernel size %d needs %d" % (taps, k, k ** 3))
    cout, cin = (b, a) if wt else (a, b)
    return ConvGeom(cin, cout, k, stride, pad, transposed, output_padding)


# ------------------------------------------------------------------------------------------------------------------
# convolution family
# ------------------------------------------------------------------------------------------------------------------
@custom_op("mra::conv3d", mutates_args=())
def conv3d(x: Tensor, w: Tensor, bias: Optional[Tensor], k: int, stride: int, pad: int, transposed: bool,
           output_padding: int, act: int, slope: float) -> Tensor:
    """nn.Conv3d / nn.ConvTranspose3d (+ bias, + fused activation) on a channels-last tensor."""
    g = _geom(w, k, stride, pad, transposed, output_padding)
    y, _ = ops.impl().conv_fprop(x.contiguous(), w.contiguous(), bias, g, act, slope, False)
    return y


@conv3d.register_fake
def _(x, w, bias, k, stride, pad, transposed, output_padding, act, slope):
    g = _geom(w, k, stride, pad, transposed, output_padding)
    return x.new_empty((x.shape[0],) + g.out_dims(tuple(x.shape[1:4])) + (g.cout,))


@custom_op("mra::conv3d_stats", mutates_args=())
def conv3d_stats(x: Tensor, w: Tensor, bias: Optional[Tensor], k: int, stride: int, pad: int, transposed: bool,
                 output_padding: int) -> Tuple[Tensor, Tensor]:
    """conv3d that also returns the InstanceNorm statistics of its output from the epilogue:
    stats[n][c] = {sum, sum of squares} (fp64) -- what ``mra::inorm_act_pad`` consumes.  Not differentiable on its own
    (the fused networks pair it with the norm's backward)."""
    g = _geom(w, k, stride, pad, transposed, output_padding)
    y, st = ops.impl().conv_fprop(x.contiguous(), w.contiguous(), bias, g, ACT_NONE, 0.0, True)
    return y, st


@conv3d_stats.register_fake
def _(x, w, bias, k, stride, pad, transposed, output_padding):
    g = _geom(w, k, stride, pad, transposed, output_padding)
    y = x.new_empty((x.shape[0],) + g.out_dims(tuple(x.shape[1:4])) + (g.cout,))
    return y, x.new_empty((x.shape[0], g.cout, 2), dtype=torch.float64)


@custom_op("mra::conv3d_dgrad", mutates_args=())
def conv3d_dgrad(dy: Tensor, wT: Tensor, in_dims: List[int], k: int, stride: int, pad: int, transposed: bool,
                 output_padding: int) -> Tensor:
    """gradient with respect to the input; ``wT`` = packed weights transposed to [taps][Cin][Cout]."""
    g = _geom(wT, k, stride, pad, transposed, output_padding, wt=True)
    return ops.impl().conv_dgrad(dy.contiguous(), wT.contiguous(), g, tuple(in_dims))


@conv3d_dgrad.register_fake
def _(dy, wT, in_dims, k, stride, pad, transposed, output_padding):
    return dy.new_empty((dy.shape[0],) + tuple(in_dims) + (wT.shape[1],))


@custom_op("mra::conv3d_wgrad", mutates_args=())
def conv3d_wgrad(x: Tensor, dy: Tensor, k: int, stride: int, pad: int, transposed: bool,
                 output_padding: int) -> Tuple[Tensor, Tensor]:
    """gradients with respect to the packed weights ([taps][Cout][Cin]) and the bias, both fp32."""
    g = ConvGeom(x.shape[4], dy.shape[4], k, stride, pad, transposed, output_padding)
    dw, db = ops.impl().conv_wgrad(x.contiguous(), dy.contiguous(), g, want_bias=True)
    return dw, db


@conv3d_wgrad.register_fake
def _(x, dy, k, stride, pad, transposed, output_padding):
    return (x.new_empty((k ** 3, dy.shape[4], x.shape[4]), dtype=torch.float32),
            x.new_empty((dy.shape[4],), dtype=torch.float32))


@custom_op("mra::act_bwd", mutates_args=())
def act_bwd(dy: Tensor, y: Tensor, act: int, slope: float) -> Tensor:
    """dy * act'(.) expressed through the activation's OUTPUT y (in-place activations keep only that)."""
    return ops.impl().act_bwd(dy.contiguous(), y.contiguous(), act, slope)


@act_bwd.register_fake
def _(dy, y, act, slope):
    return torch.empty_like(dy)


def _conv3d_setup(ctx, inputs, output):
    x, w, bias, k, stride, pad, transposed, output_padding, act, slope = inputs
    ctx.geom = (k, stride, pad, transposed, output_padding)
    ctx.act, ctx.slope, ctx.in_dims, ctx.has_bias = act, slope, list(x.shape[1:4]), bias is not None
    ctx.save_for_backward(x, w, output if act != ACT_NONE else None)


def _conv3d_backward(ctx, gy):
    x, w, y = ctx.saved_tensors
    gy = gy.contiguous()
    if ctx.act != ACT_NONE:
        gy = torch.ops.mra.act_bwd(gy, y, ctx.act, ctx.slope)
    dx = dw = db = None
    if ctx.needs_input_grad[0]:
        dx = torch.ops.mra.conv3d_dgrad(gy, w.transpose(1, 2).contiguous(), ctx.in_dims, *ctx.geom)
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        dw32, db32 = torch.ops.mra.conv3d_wgrad(x, gy, *ctx.geom)
        dw = dw32.to(w.dtype) if ctx.needs_input_grad[1] else None
        db = db32 if ctx.has_bias and ctx.needs_input_grad[2] else None
    return dx, dw, db, None, None, None, None, None, None, None


conv3d.register_autograd(_conv3d_backward, setup_context=_conv3d_setup)


# ------------------------------------------------------------------------------------------------------------------
# InstanceNorm3d (+ ReLU / LeakyReLU, + residual add, + ReplicationPad3d) in one pass
# ------------------------------------------------------------------------------------------------------------------
@custom_op("mra::inorm_act_pad", mutates_args=())
def inorm_act_pad(x: Tensor, stats: Optional[Tensor], residual: Optional[Tensor], pad: int, act: int, slope: float,
                  res_pad: int, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """y = ReplicationPad3d(pad)(act(InstanceNorm3d(x)) + crop(residual, res_pad)); also returns mean / rstd [n][c]
    (fp32) for the backward.  ``stats`` from ``mra::conv3d_stats`` saves the statistics pass."""
    I = ops.impl()
    x = x.contiguous()
    st = stats if stats is not None else I.inorm_stats(x)
    y, mean, rstd = I.inorm_fwd(x, st, residual.contiguous() if residual is not None else None, pad, act, slope,
                                res_pad if residual is not None else -1, eps)
    return y, mean, rstd


@inorm_act_pad.register_fake
def _(x, stats, residual, pad, act, slope, res_pad, eps):
    n, d, h, w, c = x.shape
    return (x.new_empty((n, d + 2 * pad, h + 2 * pad, w + 2 * pad, c)),
            x.new_empty((n, c), dtype=torch.float32), x.new_empty((n, c), dtype=torch.float32))


@custom_op("mra::inorm_act_pad_bwd", mutates_args=())
def inorm_act_pad_bwd(gy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, pad: int, act: int, slope: float,
                      res_pad: int) -> Tuple[Tensor, Tensor]:
    """(dx, dresidual); dresidual is empty when ``res_pad`` < 0."""
    dx, dres = ops.impl().inorm_bwd(gy.contiguous(), x.contiguous(), mean, rstd, pad, act, slope, res_pad)
    return dx, (dres if dres is not None else x.new_empty((0,)))


@inorm_act_pad_bwd.register_fake
def _(gy, x, mean, rstd, pad, act, slope, res_pad):
    n, d, h, w, c = x.shape
    dres = x.new_empty((n, d + 2 * res_pad, h + 2 * res_pad, w + 2 * res_pad, c)) if res_pad >= 0 else x.new_empty((0,))
    return torch.empty_like(x), dres


def _inorm_setup(ctx, inputs, output):
    x, stats, residual, pad, act, slope, res_pad, eps = inputs
    _, mean, rstd = output
    ctx.cfg = (pad, act, slope, res_pad if residual is not None else -1)
    ctx.mark_non_differentiable(mean, rstd)
    ctx.save_for_backward(x, mean, rstd)


def _inorm_backward(ctx, gy, _gmean, _grstd):
    x, mean, rstd = ctx.saved_tensors
    dx, dres = torch.ops.mra.inorm_act_pad_bwd(gy.contiguous(), x, mean, rstd, *ctx.cfg)
    return dx, None, (dres if ctx.cfg[3] >= 0 else None), None, None, None, None, None


inorm_act_pad.register_autograd(_inorm_backward, setup_context=_inorm_setup)
