"""Autograd wiring of the fused ops.

Each ``torch.autograd.Function`` below is one unit of the fused execution plan; forward and backward
both call only ``ops.impl()`` (the C-ABI kernels).  Activations inside a network are channels-last
(N, D, H, W, C) tensors in the network's compute dtype; network inputs / outputs are converted from
/ to the reference's (N, C, D, H, W) fp32 convention at the boundary (torch view/cast plumbing).
"""
import os

import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, LOSS_BCE_CONST, LOSS_L1, LOSS_MSE_CONST


_NSTATS_ALWAYS = False


def weight_grad_view(dw_packed, k, transposed):
    """[taps][Cout][Cin] fp32 -> the parameter's logical shape with the packed memory layout
    (so autograd's AccumulateGrad can adopt it without a copy)."""
    t, co, ci = dw_packed.shape
    v = dw_packed.view(k, k, k, co, ci)
    return v.permute(4, 3, 0, 1, 2) if transposed else v.permute(3, 4, 0, 1, 2)


class NormBwdLink:
    """Hand-over of the norm backward's statistics from the dgrad epilogue of the conv that consumes the norm's
    output (mra_conv3d_dgrad_nstats) to that norm's backward, which then only runs its apply pass.  Created by
    run_program() for a (fused norm -> conv) pair whose norm output has no other consumer."""
    __slots__ = ("act", "slope", "sums")

    def __init__(self, act, slope):
        self.act, self.slope, self.sums = act, slope, None


def nstats_profitable(g, x):
    """Is the norm-backward statistics pass cheaper inside this conv's dgrad epilogue than as its own kernel?
    Measured per layer on B200 (tools/nstats_bench.py).  With the first build of the epilogue (spilled loop variables,
    un-prefetched aux loads: profiles/r02_nstats_bench_v1.txt) only layers whose tiles accumulate over K = taps x Cout
    >= ~1500 won; since the epilogue warps have their own register budget and fetch the aux values ahead of use it
    wins on every linked layer (profiles/r02_nstats_bench_v3.txt: G.d1 +174 us against a 242 us statistics kernel,
    G.rb +10 vs 31, G.u2 +25 vs 69, D.3 / D.4 / D.5 +8..10 vs 23..27) except the head lowering (Cout = 1: the dual-plane
    column kernel is bound by its epilogue, +338 us vs 301).  MRA_NORM_BWD_FUSED=3 restores the rule of the first build
    for the ablation."""
    if min(g.cin, g.cout) == 1:
        return False                                       # channel-expanded lowering (column kernel)
    if os.environ.get("MRA_NORM_BWD_FUSED", "1") != "3":
        return True
    if not g.transposed and g.stride > 1:
        k_eff = (g.k ** 3) / float(g.stride ** 3) * g.cout  # dgrad of a strided conv: taps split over stride^3 phases
    else:
        k_eff = g.k ** 3 * g.cout
    return k_eff >= 1500 or x.numel() <= (1 << 23)


class ConvFn(torch.autograd.Function):
    """nn.Conv3d / nn.ConvTranspose3d (+bias, + epilogue activation, + InstanceNorm statistics)."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod, act, slope, want_stats, bias_grad, link=None):
        I = ops.impl()
        g = mod.geom
        wc = mod.packed_weight(x.dtype)
        # channel-expanded lowerings (stem / head / PatchGAN layer 0): one workspace shared by the layer's calls so
        # that the expanded operand is built once (fprop -> wgrad, or wgrad -> dgrad)
        low, ws = 0, None
        if x.dtype == torch.bfloat16 and min(g.cin, g.cout) == 1 and hasattr(I, "conv_shared_workspace"):
            low, ws = I.conv_shared_workspace(g, x.shape[0], tuple(x.shape[1:4]), x.dtype, x.device)
        keep = low in (1, 3) and ctx.needs_input_grad[1]
        if keep:
            y, stats = I.conv_fprop(x, wc, bias.detach() if bias is not None else None, g, act, slope, want_stats, ws=ws)
        else:
            y, stats = I.conv_fprop(x, wc, bias.detach() if bias is not None else None, g, act, slope, want_stats)
        ctx.low, ctx.ws = low, (ws if keep else None)
        ctx.mod, ctx.act, ctx.slope, ctx.bias_grad = mod, act, slope, bias_grad
        if link is not None and not (x.dtype == torch.bfloat16 and hasattr(I, "conv_dgrad_nstats") and
                                     (_NSTATS_ALWAYS or nstats_profitable(g, x)) and
                                     I.conv_dgrad_nstats_supported(g, x.shape[0], tuple(x.shape[1:4]), x.dtype)):
            link = None
        ctx.link = link
        ctx.in_dims = tuple(x.shape[1:4])
        ctx.save_for_backward(x, y if act != ACT_NONE else None)
        # the statistics output never receives a gradient: without this autograd hands backward() a freshly zero-filled
        # fp64 tensor for it on every call (168 fill kernels per step)
        ctx.set_materialize_grads(False)
        if stats is not None:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    def backward(ctx, gy, _gstats):
        if gy is None:                                   # (materialize_grads is off) no gradient reached the output
            return (None,) * 9
        I = ops.impl()
        x, y = ctx.saved_tensors
        mod, g = ctx.mod, ctx.mod.geom
        gy = gy.contiguous()
        if ctx.act != ACT_NONE:
            gy = I.act_bwd(gy, y, ctx.act, ctx.slope)
        dx = dw = db = None
        has_bias = mod.bias is not None
        ws2 = None
        if ctx.needs_input_grad[1]:
            want_b = has_bias and ctx.bias_grad
            # Fused gradient accumulation (single-process training).  A weight used by several passes of a step (each
            # generator runs three times) would get one fresh gradient buffer + memset per use, summed by autograd
            # with add_ kernels before AccumulateGrad runs once.  Instead this function owns ``weight.grad``: the
            # first use of a step adopts the fresh buffer, every later use lets the wgrad kernel ADD into it
            # (MRA_CONV_ACCUMULATE), and autograd is handed no gradient for the parameters at all.
            fuse = bool(getattr(mod, "fuse_wgrad", False))
            acc = {}
            if fuse:
                wg = mod.weight.grad
                if (wg is not None and wg.dtype == torch.float32 and wg.stride() == mod.weight.stride()
                        and (not want_b or mod.bias.grad is not None)):
                    acc = dict(acc_dw=mod._packed_view(wg), acc_db=mod.bias.grad if want_b else None)
                elif wg is not None:
                    fuse = False                         # foreign gradient layout: leave this one to autograd
            if ctx.ws is not None:                       # lowerings 1 / 3: the forward pass left the expanded x in ctx.ws
                dwp, dbv = I.conv_wgrad(x, gy, g, want_bias=want_b, ws=ctx.ws, reuse=True, **acc)
                ctx.ws = None
            elif ctx.low == 2 and ctx.needs_input_grad[0]:   # head: wgrad expands dy, dgrad reuses it
                _, ws2 = I.conv_shared_workspace(g, x.shape[0], ctx.in_dims, x.dtype, x.device)
                dwp, dbv = I.conv_wgrad(x, gy, g, want_bias=want_b, ws=ws2, **acc)
            else:
                dwp, dbv = I.conv_wgrad(x, gy, g, want_bias=want_b, **acc)
            if not fuse:
                dw = weight_grad_view(dwp, g.k, g.transposed)
                if has_bias and ctx.needs_input_grad[2]:
                    db = dbv if dbv is not None else torch.zeros_like(mod.bias)
            elif not acc:                                # first use of the step: adopt the fresh buffers
                mod.weight.grad = weight_grad_view(dwp, g.k, g.transposed)
                if has_bias and mod.bias.requires_grad and mod.bias.grad is None:
                    mod.bias.grad = dbv if dbv is not None else torch.zeros_like(mod.bias)
        if ctx.needs_input_grad[0]:
            kw = dict(ws=ws2, reuse=True) if ws2 is not None else {}
            if ctx.link is not None:
                # x is the fused norm's stored output: the epilogue also produces that norm's backward statistics
                dx, ctx.link.sums = I.conv_dgrad_nstats(gy, mod.packed_weight_t(gy.dtype), g, ctx.in_dims, x, ctx.link.act,
                                                        ctx.link.slope, **kw)
            else:
                dx = I.conv_dgrad(gy, mod.packed_weight_t(gy.dtype), g, ctx.in_dims, **kw)
        return dx, dw, db, None, None, None, None, None, None


class NormActPadFn(torch.autograd.Function):
    """InstanceNorm3d -> ReLU/LeakyReLU -> (+ residual) -> ReplicationPad3d, one kernel each way."""

    @staticmethod
    def forward(ctx, x, stats, residual, mod, act, slope, pad, res_pad, link=None):
        I = ops.impl()
        use_running = not mod.training and mod.track_running_stats
        ctx.link = link if not use_running else None
        if link is not None and use_running:
            link.act = None                        # tells the consumer conv not to bother (ConvFn checks sums only)
        if stats is None and not use_running:
            stats = I.inorm_stats(x)
        update = mod.training and mod.track_running_stats
        y, mean, rstd = I.inorm_fwd(x, stats, residual, pad, act, slope, res_pad if residual is not None else -1,
                                    mod.eps, mod.momentum,
                                    mod.running_mean if (update or use_running) else None,
                                    mod.running_var if (update or use_running) else None, use_running)
        ctx.cfg = (act, slope, pad, res_pad if residual is not None else -1, use_running)
        ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        I = ops.impl()
        x, mean, rstd = ctx.saved_tensors
        act, slope, pad, res_pad, use_running = ctx.cfg
        want_res = res_pad >= 0 and ctx.needs_input_grad[2]
        link = ctx.link
        if link is not None and link.sums is not None:
            sums, link.sums = link.sums, None      # statistics came out of the consumer conv's dgrad epilogue
            dx, dres = I.inorm_bwd_apply(gy.contiguous(), x, mean, rstd, sums, pad, act, slope, res_pad if want_res else -1)
        else:
            dx, dres = I.inorm_bwd(gy.contiguous(), x, mean, rstd, pad, act, slope, res_pad if want_res else -1, use_running)
        return dx, None, dres, None, None, None, None, None, None


class BatchNormActPadFn(torch.autograd.Function):
    """BatchNorm3d(affine) -> ReLU/LeakyReLU -> (+ residual) -> ReplicationPad3d on the InstanceNorm kernels.

    Batch statistics = the conv epilogue's per-sample {sum, sum of squares} added over the batch.  With
    rstd' = gamma * rstd_b and mean' = mean_b - beta / rstd' the kernels' (x - mean') * rstd' IS gamma * xhat + beta, so the
    forward is the same single pass.  Backward: the kernels' statistics pass yields sum(dy) and sum(dy * z)
    (z = gamma xhat + beta) per sample; added over the batch they give dbeta and dgamma, and
    dx = rstd' (dy - m1' - z m2') with m1' = M1 - beta M2 / gamma, m2' = M2 / gamma (M1, M2 the batch means of dy and
    dy * xhat) is exactly what the apply pass computes from the sums it is handed.  The per-channel algebra in between
    is a handful of [C]-sized torch operations."""

    @staticmethod
    def forward(ctx, x, stats, residual, weight, bias, mod, act, slope, pad, res_pad):
        I = ops.impl()
        n, d, h, w, c = x.shape
        nv = float(n * d * h * w)
        training = mod.training or not mod.track_running_stats
        if training:
            if stats is None:
                stats = I.inorm_stats(x)
            S = stats.sum(0)                                           # [C][2] fp64
            mean_b = S[:, 0] / nv
            var_b = (S[:, 1] / nv - mean_b * mean_b).clamp_min(0.0)
            if mod.track_running_stats:
                with torch.no_grad():
                    mom = mod.momentum
                    mod.running_mean.mul_(1 - mom).add_(mom * mean_b.to(mod.running_mean.dtype))
                    mod.running_var.mul_(1 - mom).add_(mom * (var_b * (nv / max(nv - 1.0, 1.0))).to(mod.running_var.dtype))
                    mod.num_batches_tracked += 1
        else:
            mean_b, var_b = mod.running_mean.double(), mod.running_var.double()
        rstd_b = 1.0 / torch.sqrt(var_b + mod.eps)
        g = weight.detach().double() if weight is not None else torch.ones_like(rstd_b)
        b = bias.detach().double() if bias is not None else torch.zeros_like(rstd_b)
        g = torch.where(g == 0, torch.full_like(g, 1e-12), g)          # keeps the folded form finite
        r_f = g * rstd_b
        m_f = mean_b - b / r_f
        mean_in = m_f.float().unsqueeze(0).expand(n, c).contiguous()
        rstd_in = r_f.float().unsqueeze(0).expand(n, c).contiguous()
        rp = res_pad if residual is not None else -1
        y, _, _ = I.inorm_fwd(x, stats, residual, pad, act, slope, rp, mod.eps, mod.momentum, given=(mean_in, rstd_in))
        ctx.cfg = (act, slope, pad, rp, training, nv)
        ctx.save_for_backward(x, mean_in, rstd_in, g, b)
        return y

    @staticmethod
    def backward(ctx, gy):
        I = ops.impl()
        x, mean_in, rstd_in, g, b = ctx.saved_tensors
        act, slope, pad, res_pad, training, nv = ctx.cfg
        n, d, h, w, c = x.shape
        v = float(d * h * w)
        want_res = res_pad >= 0 and ctx.needs_input_grad[2]
        rp = res_pad if want_res else -1
        gy = gy.contiguous()
        sums = I.inorm_bwd_stats(gy, x, mean_in, rstd_in, pad, act, slope, rp)
        T = sums.sum(0)                                                # [C][2]: sum dy, sum dy * z
        dbeta = T[:, 0]
        dgamma = (T[:, 1] - b * T[:, 0]) / g
        if training:
            M1, M2 = dbeta / nv, dgamma / nv
            m1, m2 = M1 - b * M2 / g, M2 / g
        else:
            m1, m2 = torch.zeros_like(dbeta), torch.zeros_like(dbeta)
        sums_in = (torch.stack([m1, m2], -1) * v).unsqueeze(0).expand(n, c, 2).contiguous()
        dx, dres = I.inorm_bwd_apply(gy, x, mean_in, rstd_in, sums_in, pad, act, slope, rp)
        dw = dgamma.float() if ctx.needs_input_grad[3] else None
        db = dbeta.float() if ctx.needs_input_grad[4] else None
        return dx, None, dres, dw, db, None, None, None, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) in training mode.  The keep mask is drawn from torch's generator of the tensor's device (the
    reference's masks come from the same generator, in NCDHW order: the statistics agree, the bits cannot)."""

    @staticmethod
    def forward(ctx, x, p):
        keep = torch.empty(x.shape, dtype=torch.uint8, device=x.device).bernoulli_(1.0 - p)
        ctx.scale = 1.0 / (1.0 - p)
        ctx.save_for_backward(keep)
        return ops.impl().mask_scale(x.contiguous(), keep, ctx.scale)

    @staticmethod
    def backward(ctx, gy):
        (keep,) = ctx.saved_tensors
        return ops.impl().mask_scale(gy.contiguous(), keep, ctx.scale), None


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, slope):
        y = ops.impl().act_fwd(x.contiguous(), act, slope)
        ctx.cfg = (act, slope)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return ops.impl().act_bwd(gy.contiguous(), y, *ctx.cfg), None, None


class CatActFn(torch.autograd.Function):
    """act(torch.cat([a, b], channel axis)) in one pass (the UNet skip concat + the parent block's ReLU)."""

    @staticmethod
    def forward(ctx, a, b, act, slope):
        out = ops.impl().cat2_act_fwd(a.contiguous(), b.contiguous(), act, slope)
        ctx.cfg = (a.shape[4], act, slope)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, gout):
        (out,) = ctx.saved_tensors
        ca, act, slope = ctx.cfg
        da, db = ops.impl().cat2_act_bwd(gout.contiguous(), out, ca, act, slope, (ctx.needs_input_grad[0], ctx.needs_input_grad[1]))
        return da, db, None, None


class RepPadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        ctx.pad = pad
        return ops.impl().reppad_fwd(x.contiguous(), pad)

    @staticmethod
    def backward(ctx, gy):
        return ops.impl().reppad_bwd(gy.contiguous(), ctx.pad), None


class PairLossFn(torch.autograd.Function):
    """mean |a - b| (torch.nn.L1Loss).  ``b`` is treated as a constant (it is always a real image)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return ops.impl().loss_fwd(LOSS_L1, a, b)

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        return ops.impl().loss_bwd(LOSS_L1, a, b, 0.0, gout.float().reshape(1).contiguous(), 1.0 / a.numel()), None


class ConstLossFn(torch.autograd.Function):
    """MSELoss / BCELoss against a constant label expanded to the input's shape (GANLoss)."""

    @staticmethod
    def forward(ctx, a, kind, target):
        a = a.contiguous()
        ctx.kind, ctx.target = kind, target
        ctx.save_for_backward(a)
        return ops.impl().loss_fwd(kind, a, None, target)

    @staticmethod
    def backward(ctx, gout):
        (a,) = ctx.saved_tensors
        return ops.impl().loss_bwd(ctx.kind, a, None, ctx.target, gout.float().reshape(1).contiguous(),
                                   1.0 / a.numel()), None, None


def l1_loss(a, b):
    return PairLossFn.apply(a, b.detach())


def mse_const_loss(a, target):
    return ConstLossFn.apply(a, LOSS_MSE_CONST, float(target))


def bce_const_loss(a, target):
    return ConstLossFn.apply(a, LOSS_BCE_CONST, float(target))


def cor_coe_loss(y_pred, y_target):
    """Cor_CoeLoss = 1 - r^2 from one fused pass over both volumes (5 sums); value only -- the
    reference computes it every step but never adds it to loss_G (cycle_gan_model.py:217-223)."""
    s = ops.impl().corr_sums(y_pred.detach().contiguous(), y_target.detach().contiguous())
    n = y_pred.numel()
    sx, sy, sxy, sxx, syy = s[0], s[1], s[2], s[3], s[4]
    num = sxy - sx * sy / n
    den = torch.sqrt(sxx - sx * sx / n) * torch.sqrt(syy - sy * sy / n)
    r = num / den
    return (1 - r * r).float()


# -- boundary layout helpers (views / casts only) ---------------------------------------------
def to_channels_last(x, dtype):
    """(N, C, D, H, W) any float dtype -> (N, D, H, W, C) contiguous ``dtype``."""
    if x.shape[1] == 1:
        y = x.reshape(x.shape[0], x.shape[2], x.shape[3], x.shape[4], 1)
    else:
        y = x.permute(0, 2, 3, 4, 1)
    return y.to(dtype).contiguous()


def to_channels_first(y, dtype=torch.float32):
    """(N, D, H, W, C) -> (N, C, D, H, W) contiguous ``dtype``."""
    if y.shape[4] == 1:
        x = y.reshape(y.shape[0], 1, y.shape[1], y.shape[2], y.shape[3])
    else:
        x = y.permute(0, 4, 1, 2, 3)
    return x.to(dtype).contiguous()
