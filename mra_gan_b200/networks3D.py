"""B200-native counterpart of the reference's ``models/networks3D.py``.

Same factories, class names, constructor signatures and ``state_dict`` key layout
(/root/reference/models/networks3D.py:15-118, 130-450), so checkpoints move both ways and
``train.py`` / ``test.py`` can import this module instead.  Underneath, each network compiles its
``nn.Sequential`` into a short fused program (conv -> [norm+act+residual+pad]) executed by the
hand-written sm_100a kernels through ``mra_gan_b200.ops``; there is no torch.nn arithmetic and no
CPU fallback.

Reference behaviours kept on purpose: ``padding_type='reflect'`` builds
*replication* padding (:232-235); ``use_bias`` follows the reference's InstanceNorm3d / InstanceNorm2d
comparisons (:180-183, 298-301, 384-387); the UNet skip carries LeakyReLU(x) because of the in-place
activation (:306,343); ``get_scheduler`` returns (does not raise) NotImplementedError (:40).
Added: ``unet_128`` -- the 7-down UNet the reference intended but shadowed with a duplicate
``resnet_9blocks`` branch (:94).
"""
import functools
import math
import os

import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import functional as MF
from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ConvGeom

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")  # mirrors networks3D.py:8

# compute dtype of newly built networks: bf16 = tcgen05 tensor-core path, fp32 = CUDA-core parity path
_default_compute_dtype = torch.bfloat16


def set_default_compute_dtype(dtype):
    global _default_compute_dtype
    assert dtype in (torch.float32, torch.bfloat16)
    _default_compute_dtype = dtype


def get_default_compute_dtype():
    return _default_compute_dtype


# The fused optimiser updates weights through raw pointers (no torch in-place op, so ``_version`` does not move):
# it bumps ``p._mra_epoch`` of every parameter it touched so cached compute-dtype copies refresh.  The counter is
# per parameter on purpose: a step of the discriminators' optimiser must not invalidate the generators' caches.
class _WeightsEpoch:
    value = 0          # total optimiser steps (diagnostics only)


def set_fused_wgrad(net, flag=True):
    """Let the wgrad kernels of ``net`` accumulate straight into ``weight.grad`` when a weight is used several times
    per step (functional.ConvFn.backward).  Under data parallelism (parallel.attach) it stays on by default -- the
    kernels add into the gradient-bucket views and the bucket hooks still fire once per parameter (measured: 0 late
    bucket launches); MRA_DP_FUSED_WGRAD=0 / attach(fused_wgrad=False) hands every gradient to autograd instead."""
    for m in net.modules():
        if isinstance(m, _ConvNd):
            m.fuse_wgrad = bool(flag)
    return net


###############################################################################
# Layers (parameter containers + plan tokens; arithmetic lives in the kernels)
###############################################################################
class _ConvNd(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.output_padding = kernel_size, stride, padding, output_padding
        k = kernel_size
        # physical layout [kD][kH][kW][Cout][Cin] (the kernels' packed layout), logical shape = torch's
        base = torch.empty(k, k, k, out_channels, in_channels)
        perm = (4, 3, 0, 1, 2) if self.transposed else (3, 4, 0, 1, 2)
        self.weight = nn.Parameter(base.permute(*perm))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.geom = ConvGeom(in_channels, out_channels, k, stride, padding, self.transposed, output_padding)
        self._cache = {}
        self.reset_parameters()

    def reset_parameters(self):
        # torch's default Conv init (kaiming_uniform(a=sqrt(5)) + uniform bias), drawn in torch's
        # logical element order so the host RNG stream matches nn.Conv3d / nn.ConvTranspose3d
        ref_shape = tuple(self.weight.shape)
        w = torch.empty(ref_shape)
        init.kaiming_uniform_(w, a=math.sqrt(5))
        with torch.no_grad():
            self.weight.copy_(w)
            if self.bias is not None:
                fan_in = ref_shape[1] * self.kernel_size ** 3
                bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
                init.uniform_(self.bias, -bound, bound)

    # -- packed views / compute-dtype copies -------------------------------------------------
    fuse_wgrad = False       # set per instance by set_fused_wgrad(): wgrad kernels add into weight.grad (functional.ConvFn)

    def _packed_view(self, t):
        perm = (2, 3, 4, 1, 0) if self.transposed else (2, 3, 4, 0, 1)
        p = t.permute(*perm)
        if not p.is_contiguous():            # someone replaced the parameter with a plain-layout tensor
            p = p.contiguous()
        return p.view(self.kernel_size ** 3, self.out_channels, self.in_channels)

    def _packed_master(self):
        return self._packed_view(self.weight.detach())

    def _tag(self):
        return (self.weight._version, getattr(self.weight, "_mra_epoch", 0), self.weight.data_ptr())

    def packed_weight(self, dtype):
        """[taps][Cout][Cin] in the compute dtype (the master itself for fp32)."""
        master = self._packed_master()
        if dtype == torch.float32 and master.dtype == torch.float32:
            return master
        key = ("w", dtype)
        ent = self._cache.get(key)
        tag = self._tag()
        if ent is None or ent[0] != tag:
            shadow = getattr(self.weight, "_mra_shadow", None)
            if (shadow is not None and shadow.dtype == dtype and
                    getattr(self.weight, "_mra_shadow_tag", None) == tag):
                wc = self._packed_view(shadow)                                  # written by Adam
            else:
                wc = ops.impl().convert(master, dtype)
            ent = (tag, wc)
            self._cache[key] = ent
        return ent[1]

    def packed_weight_t(self, dtype):
        """[taps][Cin][Cout] in the compute dtype (dgrad operand)."""
        key = ("wT", dtype)
        ent = self._cache.get(key)
        tag = self._tag()
        if ent is None or ent[0] != tag:
            ent = (tag, ops.impl().pack_weight_t(self._packed_master(), dtype))
            self._cache[key] = ent
        return ent[1]

    def make_shadow(self, dtype):
        """bf16 twin of the weight (same physical element order) that the fused Adam keeps current."""
        if dtype == torch.float32:
            return None
        sh = getattr(self.weight, "_mra_shadow", None)
        if sh is None or sh.dtype != dtype or sh.device != self.weight.device:
            sh = torch.empty_like(self.weight, dtype=dtype)     # preserve_format keeps the packed strides
            self.weight._mra_shadow = sh
            self.weight._mra_shadow_tag = None
        return sh

    def _apply(self, fn, *a, **k):
        self._cache = {}
        return super()._apply(fn, *a, **k)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d, padding=%d%s%s" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
            ", output_padding=%d" % self.output_padding if self.transposed else "",
            "" if self.bias is not None else ", bias=False")


class Conv3d(_ConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, 0, bias)


class ConvTranspose3d(_ConvNd):
    transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, output_padding, bias)


class InstanceNorm3d(nn.Module):
    """nn.InstanceNorm3d(affine=False, track_running_stats=True) state (networks3D.py:19)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=False, track_running_stats=False):
        super().__init__()
        if affine:
            raise NotImplementedError("affine InstanceNorm3d is not used by the reference")
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features))
            self.register_buffer("running_var", torch.ones(num_features))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.running_mean = self.running_var = self.num_batches_tracked = None

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g, affine=False, track_running_stats=%s" % (
            self.num_features, self.eps, self.momentum, self.track_running_stats)


class BatchNorm3d(nn.Module):
    """nn.BatchNorm3d(affine=True, track_running_stats=True) state (networks3D.py:17, norm='batch').  The arithmetic
    runs on the InstanceNorm kernels: batch statistics are the per-sample statistics summed over the batch, and the
    affine is folded into the (mean, rstd) pair the kernels are given (functional.BatchNormActPadFn)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features))
            self.register_buffer("running_var", torch.ones(num_features))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.running_mean = self.running_var = self.num_batches_tracked = None

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g, affine=%s, track_running_stats=%s" % (
            self.num_features, self.eps, self.momentum, self.affine, self.track_running_stats)


_NORMS = (InstanceNorm3d, BatchNorm3d)


def apply_norm(x, stats, res, m, act, slope, pad, res_pad, link=None):
    """One fused norm instruction (normalise -> activation -> + residual -> replication pad) for either norm type."""
    if isinstance(m, BatchNorm3d):
        return MF.BatchNormActPadFn.apply(x, stats, res, m.weight, m.bias, m, act, slope, pad, res_pad)
    return MF.NormActPadFn.apply(x, stats, res, m, act, slope, pad, res_pad, link)


class ReplicationPad3d(nn.Module):
    def __init__(self, padding):
        super().__init__()
        self.padding = padding


class ReLU(nn.Module):
    act, slope = ACT_RELU, 0.0

    def __init__(self, inplace=False):
        super().__init__()


class LeakyReLU(nn.Module):
    act = ACT_LRELU

    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.slope = negative_slope


class Tanh(nn.Module):
    act, slope = ACT_TANH, 0.0


class Sigmoid(nn.Module):
    act, slope = ACT_SIGMOID, 0.0


class Dropout(nn.Module):
    def __init__(self, p=0.5):
        super().__init__()
        self.p = p


_ACTS = (ReLU, LeakyReLU, Tanh, Sigmoid)


###############################################################################
# Fused-program compiler / executor for sequential nets
###############################################################################
def _flatten(seq, out):
    for m in seq:
        if isinstance(m, ResnetBlock):
            out.append(("block_begin", m))
            _flatten(m.conv_block, out)
            out.append(("block_end", m))
        elif isinstance(m, nn.Sequential):
            _flatten(m, out)
        else:
            out.append(("mod", m))
    return out


def compile_program(seq):
    """nn.Sequential -> list of fused instructions:
         ("pad", p) | ("conv", mod, act, slope, want_stats) | ("norm", mod, act, slope, pad, use_res)
         | ("act", act, slope) | ("save",) ; residual bookkeeping is positional (one live slot)."""
    toks = _flatten(seq, [])
    prog, i, halo = [], 0, 0          # halo = materialised replication halo of the current tensor
    n = len(toks)

    def peek_mod(j, cls):
        return j < n and toks[j][0] == "mod" and isinstance(toks[j][1], cls)

    while i < n:
        kind, m = toks[i]
        if kind == "block_begin":
            prog.append(("save", halo))
            i += 1
        elif kind == "block_end":
            raise NotImplementedError("a ResnetBlock must end in a normalisation layer")
        elif isinstance(m, ReplicationPad3d):
            if halo != m.padding:
                assert halo == 0
                prog.append(("pad", m.padding))
            halo = 0                   # consumed by the next conv
            i += 1
        elif isinstance(m, _ConvNd):
            j = i + 1
            if peek_mod(j, _NORMS):
                norm = toks[j][1]
                prog.append(("conv", m, ACT_NONE, 0.0, True))
                j += 1
                act, slope = ACT_NONE, 0.0
                if peek_mod(j, (ReLU, LeakyReLU)):
                    act, slope = toks[j][1].act, toks[j][1].slope
                    j += 1
                drop = None
                if peek_mod(j, Dropout):                 # ResnetBlock: [.., norm, ReLU, Dropout(0.5), pad, conv, norm]
                    drop = toks[j][1]
                    j += 1
                use_res = False
                if j < n and toks[j][0] == "block_end":
                    use_res = True
                    j += 1
                # fold the replication pad that feeds the next conv (possibly across a block boundary)
                jj, pad = j, 0
                if jj < n and toks[jj][0] == "block_begin":
                    jj += 1
                if peek_mod(jj, ReplicationPad3d) and drop is None:
                    pad = toks[jj][1].padding            # (a dropout in between keeps the pad a separate instruction)
                prog.append(("norm", norm, act, slope, pad, use_res))
                if drop is not None:
                    prog.append(("dropout", drop))
                halo = pad
                i = j
            else:
                act, slope = ACT_NONE, 0.0
                if peek_mod(j, _ACTS):
                    act, slope = toks[j][1].act, toks[j][1].slope
                    j += 1
                prog.append(("conv", m, act, slope, False))
                halo = 0
                i = j
        elif isinstance(m, _ACTS):
            prog.append(("act", m.act, m.slope))
            i += 1
        elif isinstance(m, _NORMS):
            prog.append(("norm", m, ACT_NONE, 0.0, 0, False))
            i += 1
        elif isinstance(m, Dropout):
            prog.append(("dropout", m))
            i += 1
        else:
            raise NotImplementedError("layer %s is not supported by the fused executor" % type(m).__name__)
    return prog


# Norm-backward statistics from the consumer conv's dgrad epilogue (functional.NormBwdLink): on by default for the
# bf16 path, MRA_NORM_BWD_FUSED=0 restores the two-pass norm backward everywhere (ablation).
_FUSE_NORM_BWD = os.environ.get("MRA_NORM_BWD_FUSED", "1") != "0"


def run_program(prog, x):
    """x: channels-last activation.  Returns the channels-last output."""
    stats = None
    saved, saved_pad = None, 0
    link = None
    for i, ins in enumerate(prog):
        op = ins[0]
        if op == "conv":
            _, m, act, slope, want_stats = ins
            # a conv followed by a train-mode InstanceNorm has a dead bias gradient
            x, stats = MF.ConvFn.apply(x, m.weight, m.bias, m, act, slope, want_stats, not want_stats,
                                       link if (link is not None and link.act is not None) else None)
            link = None
        elif op == "norm":
            _, m, act, slope, pad, use_res = ins
            res = saved if use_res else None
            # the norm's output feeds the NEXT instruction only, and that is a conv: its dgrad epilogue can produce
            # this norm's backward statistics (InstanceNorm with a piecewise-linear activation, training mode, bf16; not
            # with a residual added after the activation: xhat could not be recovered from the stored output)
            link = None
            if (_FUSE_NORM_BWD and not use_res and isinstance(m, InstanceNorm3d) and act in (ACT_NONE, ACT_RELU, ACT_LRELU)
                    and 0.0 <= slope <= 1.0 and x.dtype == torch.bfloat16 and torch.is_grad_enabled()
                    and i + 1 < len(prog) and prog[i + 1][0] == "conv"):
                link = MF.NormBwdLink(act, slope)
            x = apply_norm(x, stats, res, m, act, slope, pad, saved_pad, link)
            stats = None
        elif op == "pad":
            x = MF.RepPadFn.apply(x, ins[1])
        elif op == "save":
            saved, saved_pad = x, ins[1]
        elif op == "act":
            x = MF.ActFn.apply(x, ins[1], ins[2])
        elif op == "dropout":
            if ins[1].training and ins[1].p > 0:
                x = MF.DropoutFn.apply(x, ins[1].p)
        else:
            raise AssertionError(op)
    return x


class _FusedNet(nn.Module):
    """Base of the sequential networks: owns the compute dtype, the compiled program and the
    (N,C,D,H,W) fp32 <-> channels-last boundary."""
    _seq_attr = "model"

    def _init_fused(self):
        self.compute_dtype = _default_compute_dtype
        self._program = None

    def _seq(self):
        return getattr(self, self._seq_attr)

    def program(self):
        if self._program is None:
            self._program = compile_program(self._seq())
        return self._program

    def forward(self, input):
        x = MF.to_channels_last(input, self.compute_dtype)
        y = run_program(self.program(), x)
        return MF.to_channels_first(y, torch.float32)

    def conv_modules(self):
        return [m for m in self.modules() if isinstance(m, _ConvNd)]


###############################################################################
# Helper Functions (same names / behaviour as the reference)
###############################################################################
def get_norm_layer(norm_type='instance'):
    if norm_type == 'batch':
        norm_layer = functools.partial(BatchNorm3d, affine=True)
    elif norm_type == 'instance':
        norm_layer = functools.partial(InstanceNorm3d, affine=False, track_running_stats=True)
    elif norm_type == 'none':
        norm_layer = None
    else:
        raise NotImplementedError('normalization layer [%s] is not found' % norm_type)
    return norm_layer


def get_scheduler(optimizer, opt):
    if opt.lr_policy == 'lambda':
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + 1 + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)
        scheduler = lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    elif opt.lr_policy == 'step':
        scheduler = lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    elif opt.lr_policy == 'plateau':
        scheduler = lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    elif opt.lr_policy == 'cosine':
        scheduler = lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.niter, eta_min=0)
    else:
        return NotImplementedError('learning rate policy [%s] is not implemented', opt.lr_policy)
    return scheduler


def init_weights(net, init_type='normal', gain=0.02):
    """Same rule as the reference (networks3D.py:44-65).  Values are drawn into a torch-layout
    temporary on the parameter's device, so the RNG stream equals the reference's on that device."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            w = torch.empty(tuple(m.weight.shape), device=m.weight.device, dtype=m.weight.dtype)
            if init_type == 'normal':
                init.normal_(w, 0.0, gain)
            elif init_type == 'xavier':
                init.xavier_normal_(w, gain=gain)
            elif init_type == 'kaiming':
                init.kaiming_normal_(w, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(w, gain=gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            with torch.no_grad():
                m.weight.copy_(w)
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm3d') != -1:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    print('initialize network with %s' % init_type)
    net.apply(init_func)


def init_net(net, init_type='normal', init_gain=0.02, gpu_ids=[]):
    net.to(device)
    init_weights(net, init_type, gain=init_gain)
    return net


def define_G(input_nc, output_nc, ngf, netG, norm='batch', use_dropout=False, init_type='normal', init_gain=0.02,
             gpu_ids=[]):
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == 'resnet_9blocks':
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=9)
    elif netG == 'resnet_6blocks':
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=6)
    elif netG == 'unet_custom':
        net = UnetGenerator(input_nc, output_nc, 5, ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    elif netG == 'unet_128':
        net = UnetGenerator(input_nc, output_nc, 7, ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    elif netG == 'unet_256':
        net = UnetGenerator(input_nc, output_nc, 8, ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    elif netG == 'Dynet':
        raise NotImplementedError('Dynet wraps monai.networks.nets.DynUNet, which is outside this build (SURVEY.md 8f)')
    else:
        raise NotImplementedError('Generator model name [%s] is not recognized' % netG)
    return init_net(net, init_type, init_gain, gpu_ids)


def define_D(input_nc, ndf, netD, n_layers_D=3, norm='batch', use_sigmoid=False, init_type='normal', init_gain=0.02,
             gpu_ids=[]):
    norm_layer = get_norm_layer(norm_type=norm)
    if netD == 'basic':
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    elif netD == 'n_layers':
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    elif netD == 'pixel':
        net = PixelDiscriminator(input_nc, ndf, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    else:
        raise NotImplementedError('Discriminator model name [%s] is not recognized' % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


##############################################################################
# Losses
##############################################################################
class GANLoss(nn.Module):
    """LSGAN (MSE) or vanilla (BCE on sigmoid outputs) against a constant label
    (networks3D.py:130-150); loss and its gradient come from the fused reduction kernels."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self.use_lsgan = use_lsgan
        self._labels = (float(target_real_label), float(target_fake_label))

    def get_target_tensor(self, input, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(input)

    def __call__(self, input, target_is_real):
        target = self._labels[0] if target_is_real else self._labels[1]
        return MF.mse_const_loss(input, target) if self.use_lsgan else MF.bce_const_loss(input, target)


class L1Loss(nn.Module):
    """torch.nn.L1Loss() as used for the cycle / identity terms (cycle_gan_model.py:104-105)."""

    def forward(self, input, target):
        return MF.l1_loss(input, target)


def Cor_CoeLoss(y_pred, y_target):
    return MF.cor_coe_loss(y_pred, y_target)


##############################################################################
# Networks
##############################################################################
class ResnetGenerator(_FusedNet):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=InstanceNorm3d, use_dropout=False, n_blocks=6,
                 padding_type='reflect'):
        assert n_blocks >= 0
        super().__init__()
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        use_bias = (norm_layer.func if type(norm_layer) == functools.partial else norm_layer) == InstanceNorm3d
        model = [ReplicationPad3d(3), Conv3d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias),
                 norm_layer(ngf), ReLU(True)]
        n_downsampling = 2
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [Conv3d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                      norm_layer(ngf * mult * 2), ReLU(True)]
        mult = 2 ** n_downsampling
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer,
                                  use_dropout=use_dropout, use_bias=use_bias)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [ConvTranspose3d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                      output_padding=1, bias=use_bias),
                      norm_layer(int(ngf * mult / 2)), ReLU(True)]
        model += [ReplicationPad3d(3)]
        model += [Conv3d(ngf, output_nc, kernel_size=7, padding=0)]
        model += [Tanh()]
        self.model = nn.Sequential(*model)
        self._init_fused()


class ResnetBlock(nn.Module):
    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        conv_block = []
        for half in range(2):
            p = 0
            if padding_type in ('reflect', 'replicate'):      # both are ReplicationPad3d in the reference
                conv_block += [ReplicationPad3d(1)]
            elif padding_type == 'zero':
                p = 1
            else:
                raise NotImplementedError('padding [%s] is not implemented' % padding_type)
            conv_block += [Conv3d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim)]
            if half == 0:
                conv_block += [ReLU(True)]
                if use_dropout:
                    conv_block += [Dropout(0.5)]
        return nn.Sequential(*conv_block)

    def forward(self, x):
        raise RuntimeError("ResnetBlock is executed by its parent network's fused program")


class NLayerDiscriminator(_FusedNet):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=InstanceNorm3d, use_sigmoid=False):
        super().__init__()
        use_bias = (norm_layer.func if type(norm_layer) == functools.partial else norm_layer) == InstanceNorm3d
        kw, padw = 4, 1
        sequence = [Conv3d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [Conv3d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                         norm_layer(ndf * nf_mult), LeakyReLU(0.2, True)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [Conv3d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                     norm_layer(ndf * nf_mult), LeakyReLU(0.2, True)]
        sequence += [Conv3d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw)]
        if use_sigmoid:
            sequence += [Sigmoid()]
        self.model = nn.Sequential(*sequence)
        self._init_fused()


class PixelDiscriminator(_FusedNet):
    _seq_attr = "net"

    def __init__(self, input_nc, ndf=64, norm_layer=InstanceNorm3d, use_sigmoid=False):
        super().__init__()
        use_bias = (norm_layer.func if type(norm_layer) == functools.partial else norm_layer) == InstanceNorm3d
        net = [Conv3d(input_nc, ndf, kernel_size=1, stride=1, padding=0), LeakyReLU(0.2, True),
               Conv3d(ndf, ndf * 2, kernel_size=1, stride=1, padding=0, bias=use_bias), norm_layer(ndf * 2),
               LeakyReLU(0.2, True), Conv3d(ndf * 2, 1, kernel_size=1, stride=1, padding=0, bias=use_bias)]
        if use_sigmoid:
            net.append(Sigmoid())
        self.net = nn.Sequential(*net)
        self._init_fused()


class UnetGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=InstanceNorm3d, use_dropout=False):
        super().__init__()
        unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=None, norm_layer=norm_layer,
                                             innermost=True)
        for i in range(num_downs - 5):
            unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=unet_block,
                                                 norm_layer=norm_layer, use_dropout=use_dropout)
        unet_block = UnetSkipConnectionBlock(ngf * 4, ngf * 8, input_nc=None, submodule=unet_block, norm_layer=norm_layer)
        unet_block = UnetSkipConnectionBlock(ngf * 2, ngf * 4, input_nc=None, submodule=unet_block, norm_layer=norm_layer)
        unet_block = UnetSkipConnectionBlock(ngf, ngf * 2, input_nc=None, submodule=unet_block, norm_layer=norm_layer)
        unet_block = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=unet_block, outermost=True,
                                             norm_layer=norm_layer)
        self.model = unet_block
        self.compute_dtype = _default_compute_dtype

    def forward(self, input):
        x = MF.to_channels_last(input, self.compute_dtype)
        return MF.to_channels_first(self.model.run(x), torch.float32)

    def conv_modules(self):
        return [m for m in self.modules() if isinstance(m, _ConvNd)]


class UnetSkipConnectionBlock(nn.Module):
    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=InstanceNorm3d, use_dropout=False):
        super().__init__()
        self.outermost, self.innermost = outermost, innermost
        # the reference compares against nn.InstanceNorm2d here, so use_bias is False for the 3-D norm
        use_bias = False
        if input_nc is None:
            input_nc = outer_nc
        downconv = Conv3d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        downrelu, downnorm = LeakyReLU(0.2, True), norm_layer(inner_nc)
        uprelu, upnorm = ReLU(True), norm_layer(outer_nc)
        if outermost:
            upconv = ConvTranspose3d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv, submodule, uprelu, upconv, Tanh()]
        elif innermost:
            upconv = ConvTranspose3d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv, uprelu, upconv, upnorm]
        else:
            upconv = ConvTranspose3d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv, downnorm, submodule, uprelu, upconv, upnorm]
            if use_dropout:
                model = model + [Dropout(0.5)]          # networks3D.py:332-333
        self.model = nn.Sequential(*model)

    def run(self, x):
        """x: channels-last.  A non-outermost block returns the PAIR (LeakyReLU(x), up(...)) whose channel concat is the
        reference's output (networks3D.py:339-343) -- the skip carries the activated tensor (in-place quirk of the
        reference).  The only consumer of that concat is the parent's up-path ReLU(True) + ConvTranspose3d, so the
        parent writes ReLU(cat(.)) straight into one buffer (functional.CatActFn: no torch.cat, no separate ReLU)."""
        m = self.model
        if self.outermost:
            d, _ = MF.ConvFn.apply(x, m[0].weight, m[0].bias, m[0], ACT_NONE, 0.0, False, True)
            u = MF.CatActFn.apply(*m[1].run(d), ACT_RELU, 0.0)
            y, _ = MF.ConvFn.apply(u, m[3].weight, m[3].bias, m[3], ACT_TANH, 0.0, False, True)
            return y
        xs = MF.ActFn.apply(x, ACT_LRELU, m[0].slope)
        if self.innermost:
            # the ReLU that follows has no norm in between: fuse it into the conv epilogue
            d, _ = MF.ConvFn.apply(xs, m[1].weight, m[1].bias, m[1], ACT_RELU, 0.0, False, True)
            u, st = MF.ConvFn.apply(d, m[3].weight, m[3].bias, m[3], ACT_NONE, 0.0, True, False)
            u = apply_norm(u, st, None, m[4], ACT_NONE, 0.0, 0, -1)
        else:
            d, st = MF.ConvFn.apply(xs, m[1].weight, m[1].bias, m[1], ACT_NONE, 0.0, True, False)
            d = apply_norm(d, st, None, m[2], ACT_NONE, 0.0, 0, -1)
            u = MF.CatActFn.apply(*m[3].run(d), ACT_RELU, 0.0)
            u, st = MF.ConvFn.apply(u, m[5].weight, m[5].bias, m[5], ACT_NONE, 0.0, True, False)
            u = apply_norm(u, st, None, m[6], ACT_NONE, 0.0, 0, -1)
            if len(m) > 7 and m[7].training and m[7].p > 0:
                u = MF.DropoutFn.apply(u, m[7].p)
        return xs, u

    def forward(self, x):
        raise RuntimeError("UnetSkipConnectionBlock is executed through UnetGenerator.forward")
