"""Thin custom-op layer: torch tensors in, C-ABI calls (libmra_b200.so) out.

Tensors are channels-last activations shaped (N, D, H, W, C), contiguous, fp32 or bf16.  PyTorch
only supplies device memory and the current stream; every computation below is one of the
hand-written sm_100a kernels behind include/mra_gan_b200.h.  There is NO CPU implementation in this
package: calling an op on a non-CUDA tensor raises.  (The CPU test-suite injects an emulator that
lives under tests/, see ``set_impl``.)
"""
import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, LOSS_BCE_CONST, LOSS_L1,
                   LOSS_MSE_CONST, MRA_BF16, MRA_F32)

__all__ = ["ConvGeom", "ACT_NONE", "ACT_RELU", "ACT_LRELU", "ACT_TANH", "ACT_SIGMOID", "LOSS_L1",
           "LOSS_MSE_CONST", "LOSS_BCE_CONST"]


@dataclass(frozen=True)
class ConvGeom:
    """Geometry of one nn.Conv3d / nn.ConvTranspose3d (cubic kernel, isotropic stride)."""
    cin: int
    cout: int
    k: int
    stride: int = 1
    pad: int = 0
    transposed: bool = False
    output_padding: int = 0

    def out_dims(self, in_dims):
        if not self.transposed:
            return tuple((i + 2 * self.pad - self.k) // self.stride + 1 for i in in_dims)
        return tuple((i - 1) * self.stride - 2 * self.pad + self.k + self.output_padding for i in in_dims)

    @property
    def taps(self):
        return self.k ** 3


def _dt(t):
    if t.dtype == torch.float32:
        return MRA_F32
    if t.dtype == torch.bfloat16:
        return MRA_BF16
    raise TypeError("unsupported dtype %s (fp32 / bf16 only)" % t.dtype)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class CudaImpl:
    """The product path: every method launches kernels from libmra_b200.so on the current stream."""
    name = "cuda"

    def __init__(self):
        self.L = _lib.lib()
        self.force_naive = False

    # -- helpers ----------------------------------------------------------------------------
    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _need(*ts):
        for t in ts:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("mra_gan_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor"
                                   % t.device)
            if not t.is_contiguous():
                raise RuntimeError("mra_gan_b200 ops need contiguous tensors")

    def _conv_desc(self, g, n, in_dims, out_dims, dtype, act=ACT_NONE, slope=0.2, flags=0):
        d = _lib.ConvDesc()
        d.n, d.cin, d.cout = n, g.cin, g.cout
        d.din, d.hin, d.win = in_dims
        d.dout, d.hout, d.wout = out_dims
        d.k, d.stride, d.pad, d.transposed = g.k, g.stride, g.pad, int(g.transposed)
        d.dtype, d.act, d.slope = dtype, act, slope
        d.flags = flags | (_lib.CONV_FORCE_NAIVE if self.force_naive else 0)
        return d

    def _workspace(self, d, which, device, ws=None):
        nbytes = int(self.L.mra_conv3d_workspace_size(C.byref(d), which))
        if nbytes == 0:
            return None, 0
        if ws is not None:
            if ws.numel() < nbytes:
                raise RuntimeError("shared conv workspace too small")
            return ws, ws.numel()
        return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes

    def conv_shared_workspace(self, g, n, in_dims, dtype, device):
        """(lowering, workspace) for layers served by a channel-expanded lowering (conv_special.cuh): one buffer
        big enough for all three calls, so that the expanded operand written by one call (fprop for lowerings
        1 / 3, wgrad for lowering 2) can be reused by the next (``reuse=True``).  (0, None) for ordinary layers."""
        d = self._conv_desc(g, n, tuple(in_dims), g.out_dims(tuple(in_dims)), MRA_BF16 if dtype == torch.bfloat16 else MRA_F32)
        low = int(self.L.mra_conv3d_lowering(C.byref(d)))
        if low == 0:
            return 0, None
        nbytes = max(int(self.L.mra_conv3d_workspace_size(C.byref(d), w)) for w in range(3))
        return low, torch.empty(nbytes, dtype=torch.uint8, device=device)

    # -- convolution family -----------------------------------------------------------------
    def conv_fprop(self, x, w, bias, g, act=ACT_NONE, slope=0.2, want_stats=False, ws=None):
        self._need(x, w, bias)
        n, in_dims = x.shape[0], tuple(x.shape[1:4])
        assert x.shape[4] == g.cin and w.shape == (g.taps, g.cout, g.cin) and w.dtype == x.dtype
        out_dims = g.out_dims(in_dims)
        y = torch.empty((n,) + out_dims + (g.cout,), dtype=x.dtype, device=x.device)
        stats = torch.empty((n, g.cout, 2), dtype=torch.float64, device=x.device) if want_stats else None
        d = self._conv_desc(g, n, in_dims, out_dims, _dt(x), act, slope)
        ws, wsb = self._workspace(d, 0, x.device, ws)
        _lib.check(self.L.mra_conv3d_fprop(C.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(y), _ptr(stats),
                                           _ptr(ws), wsb, self._stream()), "mra_conv3d_fprop")
        return y, stats

    def conv_dgrad(self, dy, wT, g, in_dims, ws=None, reuse=False):
        self._need(dy, wT)
        n, out_dims = dy.shape[0], tuple(dy.shape[1:4])
        assert dy.shape[4] == g.cout and wT.shape == (g.taps, g.cin, g.cout) and wT.dtype == dy.dtype
        dx = torch.empty((n,) + tuple(in_dims) + (g.cin,), dtype=dy.dtype, device=dy.device)
        d = self._conv_desc(g, n, tuple(in_dims), out_dims, _dt(dy), flags=_lib.CONV_WS_REUSE if reuse else 0)
        ws, wsb = self._workspace(d, 1, dy.device, ws)
        _lib.check(self.L.mra_conv3d_dgrad(C.byref(d), _ptr(dy), _ptr(wT), _ptr(dx), _ptr(ws), wsb, self._stream()),
                   "mra_conv3d_dgrad")
        return dx

    def conv_dgrad_nstats_supported(self, g, n, in_dims, dtype):
        d = self._conv_desc(g, n, tuple(in_dims), g.out_dims(tuple(in_dims)), MRA_BF16 if dtype == torch.bfloat16 else MRA_F32)
        return bool(self.L.mra_conv3d_dgrad_nstats_supported(C.byref(d)))

    def conv_dgrad_nstats(self, dy, wT, g, in_dims, y_act, norm_act, norm_slope, ws=None, reuse=False):
        """dgrad + the statistics pass of the fused norm in front of the conv (mra_conv3d_dgrad_nstats): returns
        (dx, sums[n][cin][2] fp64) where sums is what ``inorm_bwd_stats(dx, x_norm, ...)`` would compute."""
        self._need(dy, wT, y_act)
        n, out_dims = dy.shape[0], tuple(dy.shape[1:4])
        assert dy.shape[4] == g.cout and wT.shape == (g.taps, g.cin, g.cout) and wT.dtype == dy.dtype
        assert tuple(y_act.shape) == (n,) + tuple(in_dims) + (g.cin,) and y_act.dtype == dy.dtype
        dx = torch.empty((n,) + tuple(in_dims) + (g.cin,), dtype=dy.dtype, device=dy.device)
        sums = torch.empty((n, g.cin, 2), dtype=torch.float64, device=dy.device)
        d = self._conv_desc(g, n, tuple(in_dims), out_dims, _dt(dy), flags=_lib.CONV_WS_REUSE if reuse else 0)
        ws, wsb = self._workspace(d, 1, dy.device, ws)
        _lib.check(self.L.mra_conv3d_dgrad_nstats(C.byref(d), _ptr(dy), _ptr(wT), _ptr(dx), _ptr(y_act), int(norm_act),
                                                  float(norm_slope), _ptr(sums), _ptr(ws), wsb, self._stream()),
                   "mra_conv3d_dgrad_nstats")
        return dx, sums

    def conv_wgrad(self, x, dy, g, want_bias=False, ws=None, reuse=False, acc_dw=None, acc_db=None):
        """dw [taps][Cout][Cin] fp32 (+ db).  With ``acc_dw`` (and ``acc_db`` when a bias gradient is wanted) the
        kernels ADD into those buffers (MRA_CONV_ACCUMULATE) instead of writing fresh ones."""
        self._need(x, dy)
        n, in_dims, out_dims = x.shape[0], tuple(x.shape[1:4]), tuple(dy.shape[1:4])
        flags = _lib.CONV_WS_REUSE if reuse else 0
        if acc_dw is not None:
            if want_bias and acc_db is None:
                raise ValueError("conv_wgrad: accumulate mode needs a bias-gradient buffer too")
            self._need(acc_dw)
            dw, db = acc_dw, (acc_db if want_bias else None)
            flags |= _lib.CONV_ACCUMULATE
        else:
            dw = torch.empty((g.taps, g.cout, g.cin), dtype=torch.float32, device=x.device)
            db = torch.empty((g.cout,), dtype=torch.float32, device=x.device) if want_bias else None
        d = self._conv_desc(g, n, in_dims, out_dims, _dt(x), flags=flags)
        ws, wsb = self._workspace(d, 2, x.device, ws)
        _lib.check(self.L.mra_conv3d_wgrad(C.byref(d), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), _ptr(ws), wsb,
                                           self._stream()), "mra_conv3d_wgrad")
        return dw, db

    def conv_uses_tensor_cores(self, g, n, in_dims, dtype, which):
        d = self._conv_desc(g, n, tuple(in_dims), g.out_dims(tuple(in_dims)),
                            MRA_BF16 if dtype == torch.bfloat16 else MRA_F32)
        return bool(self.L.mra_conv3d_uses_tensor_cores(C.byref(d), which))

    def pack_weight_t(self, w, dst_dtype):
        self._need(w)
        taps, cout, cin = w.shape
        wT = torch.empty((taps, cin, cout), dtype=dst_dtype, device=w.device)
        _lib.check(self.L.mra_pack_weight_t(_ptr(w), _dt(w), _ptr(wT), _dt(wT), taps, cout, cin, self._stream()),
                   "mra_pack_weight_t")
        return wT

    def convert(self, t, dst_dtype):
        self._need(t)
        if t.dtype == dst_dtype:
            return t
        out = torch.empty(t.shape, dtype=dst_dtype, device=t.device)
        _lib.check(self.L.mra_convert(_ptr(t), _dt(t), _ptr(out), _dt(out), t.numel(), self._stream()), "mra_convert")
        return out

    # -- instance norm family ---------------------------------------------------------------
    @staticmethod
    def _norm_desc(x, pad, act, slope, res_pad, eps=1e-5, momentum=0.1, use_running=False):
        d = _lib.NormDesc()
        d.n, d.d, d.h, d.w, d.c = x.shape
        d.pad, d.act, d.slope, d.res_pad, d.dtype = pad, act, slope, res_pad, _dt(x)
        d.eps, d.momentum, d.use_running = eps, momentum, int(use_running)
        return d

    def inorm_stats(self, x):
        self._need(x)
        stats = torch.empty((x.shape[0], x.shape[4], 2), dtype=torch.float64, device=x.device)
        d = self._norm_desc(x, 0, ACT_NONE, 0.0, -1)
        _lib.check(self.L.mra_inorm_stats(C.byref(d), _ptr(x), _ptr(stats), self._stream()), "mra_inorm_stats")
        return stats

    def inorm_fwd(self, x, stats, residual=None, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1, eps=1e-5,
                  momentum=0.1, running_mean=None, running_var=None, use_running=False, given=None):
        """``given=(mean, rstd)`` ([n][c] fp32): normalise with exactly these (use_running mode 2 of the C ABI; batch norm
        passes its batch statistics with the affine folded in) instead of deriving them from ``stats``."""
        self._need(x, stats, residual, running_mean, running_var)
        n, dd, hh, ww, c = x.shape
        y = torch.empty((n, dd + 2 * pad, hh + 2 * pad, ww + 2 * pad, c), dtype=x.dtype, device=x.device)
        if given is not None:
            mean, rstd = given
            self._need(mean, rstd)
            use_running = 2
        else:
            mean = torch.empty((n, c), dtype=torch.float32, device=x.device)
            rstd = torch.empty((n, c), dtype=torch.float32, device=x.device)
        d = self._norm_desc(x, pad, act, slope, res_pad if residual is not None else -1, eps, momentum, use_running)
        _lib.check(self.L.mra_inorm_act_pad_fwd(C.byref(d), _ptr(x), _ptr(stats), _ptr(residual), _ptr(y), _ptr(mean),
                                                _ptr(rstd), _ptr(running_mean), _ptr(running_var), self._stream()),
                   "mra_inorm_act_pad_fwd")
        return y, mean, rstd

    def inorm_bwd(self, gy, x, mean, rstd, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1, use_running=False):
        self._need(gy, x, mean, rstd)
        n, dd, hh, ww, c = x.shape
        assert tuple(gy.shape) == (n, dd + 2 * pad, hh + 2 * pad, ww + 2 * pad, c)
        dx = torch.empty_like(x)
        dres = None
        if res_pad >= 0:
            dres = torch.empty((n, dd + 2 * res_pad, hh + 2 * res_pad, ww + 2 * res_pad, c), dtype=x.dtype,
                               device=x.device)
        sums = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
        d = self._norm_desc(x, pad, act, slope, res_pad, use_running=use_running)
        _lib.check(self.L.mra_inorm_act_pad_bwd(C.byref(d), _ptr(gy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(dx),
                                                _ptr(dres), _ptr(sums), self._stream()), "mra_inorm_act_pad_bwd")
        return dx, dres

    def inorm_bwd_stats(self, gy, x, mean, rstd, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1):
        """First pass of inorm_bwd on its own: sums[n][c] = {sum dy, sum dy * xhat} (fp64), dy = act'(xhat) fold(gy)."""
        self._need(gy, x, mean, rstd)
        n, c = x.shape[0], x.shape[4]
        sums = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
        d = self._norm_desc(x, pad, act, slope, res_pad, use_running=2)
        _lib.check(self.L.mra_inorm_act_pad_bwd_stats(C.byref(d), _ptr(gy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(sums),
                                                      self._stream()), "mra_inorm_act_pad_bwd_stats")
        return sums

    def inorm_bwd_apply(self, gy, x, mean, rstd, sums, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1):
        """Second pass on its own: dx = rstd (dy - sums[.,0]/V - xhat sums[.,1]/V) with the caller's ``sums``."""
        self._need(gy, x, mean, rstd, sums)
        n, dd, hh, ww, c = x.shape
        dx = torch.empty_like(x)
        dres = None
        if res_pad >= 0:
            dres = torch.empty((n, dd + 2 * res_pad, hh + 2 * res_pad, ww + 2 * res_pad, c), dtype=x.dtype, device=x.device)
        d = self._norm_desc(x, pad, act, slope, res_pad, use_running=2)
        _lib.check(self.L.mra_inorm_act_pad_bwd_apply(C.byref(d), _ptr(gy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(sums),
                                                      _ptr(dx), _ptr(dres), self._stream()), "mra_inorm_act_pad_bwd_apply")
        return dx, dres

    def act_fwd(self, x, act, slope=0.2):
        self._need(x)
        y = torch.empty_like(x)
        _lib.check(self.L.mra_act_fwd(_ptr(x), _ptr(y), x.numel(), act, slope, _dt(x), self._stream()), "mra_act_fwd")
        return y

    def act_bwd(self, dy, y, act, slope=0.2):
        self._need(dy, y)
        dx = torch.empty_like(dy)
        _lib.check(self.L.mra_act_bwd(_ptr(dy), _ptr(y), _ptr(dx), y.numel(), act, slope, _dt(y), self._stream()),
                   "mra_act_bwd")
        return dx

    def cat2_act_fwd(self, a, b, act, slope=0.0):
        """out[..., :Ca] = act(a), out[..., Ca:] = act(b): the UNet skip concat with the following activation fused."""
        self._need(a, b)
        assert a.shape[:4] == b.shape[:4] and a.dtype == b.dtype
        ca, cb = a.shape[4], b.shape[4]
        out = torch.empty(tuple(a.shape[:4]) + (ca + cb,), dtype=a.dtype, device=a.device)
        _lib.check(self.L.mra_cat2_act_fwd(_ptr(a), _ptr(b), _ptr(out), a.numel() // ca, ca, cb, act, slope, _dt(a),
                                           self._stream()), "mra_cat2_act_fwd")
        return out

    def cat2_act_bwd(self, dout, out, ca, act, slope=0.0, want=(True, True)):
        self._need(dout, out)
        cb = out.shape[4] - ca
        da = torch.empty(tuple(out.shape[:4]) + (ca,), dtype=out.dtype, device=out.device) if want[0] else None
        db = torch.empty(tuple(out.shape[:4]) + (cb,), dtype=out.dtype, device=out.device) if want[1] else None
        _lib.check(self.L.mra_cat2_act_bwd(_ptr(dout), _ptr(out), _ptr(da), _ptr(db), out.numel() // (ca + cb), ca, cb, act,
                                           slope, _dt(out), self._stream()), "mra_cat2_act_bwd")
        return da, db

    def mask_scale(self, x, keep, scale):
        """y = x * keep * scale (dropout forward, and its own backward on the gradient); keep: uint8 0 / 1."""
        self._need(x, keep)
        y = torch.empty_like(x)
        _lib.check(self.L.mra_mask_scale(_ptr(x), _ptr(keep), _ptr(y), x.numel(), float(scale), _dt(x), self._stream()),
                   "mra_mask_scale")
        return y

    def reppad_fwd(self, x, pad):
        self._need(x)
        n, dd, hh, ww, c = x.shape
        y = torch.empty((n, dd + 2 * pad, hh + 2 * pad, ww + 2 * pad, c), dtype=x.dtype, device=x.device)
        _lib.check(self.L.mra_reppad_fwd(_ptr(x), _ptr(y), n, dd, hh, ww, c, pad, _dt(x), self._stream()),
                   "mra_reppad_fwd")
        return y

    def reppad_bwd(self, gy, pad):
        self._need(gy)
        n, dp, hp, wp, c = gy.shape
        dd, hh, ww = dp - 2 * pad, hp - 2 * pad, wp - 2 * pad
        dx = torch.empty((n, dd, hh, ww, c), dtype=gy.dtype, device=gy.device)
        _lib.check(self.L.mra_reppad_bwd(_ptr(gy), _ptr(dx), n, dd, hh, ww, c, pad, _dt(gy), self._stream()),
                   "mra_reppad_bwd")
        return dx

    # -- losses -----------------------------------------------------------------------------
    def loss_fwd(self, kind, a, b=None, target=0.0):
        """Mean-reduced loss as a 0-d fp32 tensor (no host sync)."""
        self._need(a, b)
        acc = torch.zeros(1, dtype=torch.float64, device=a.device)
        _lib.check(self.L.mra_loss_fwd(kind, _ptr(a), _ptr(b), float(target), a.numel(), _dt(a), _ptr(acc),
                                       self._stream()), "mra_loss_fwd")
        return (acc[0] / a.numel()).float()

    def loss_bwd(self, kind, a, b, target, gout, scale):
        """da = gout * scale * dloss_sum/da  (scale carries 1/numel and any lambda)."""
        self._need(a, b, gout)
        da = torch.empty_like(a)
        _lib.check(self.L.mra_loss_bwd(kind, _ptr(a), _ptr(b), float(target), a.numel(), _dt(a), _ptr(gout),
                                       float(scale), _ptr(da), self._stream()), "mra_loss_bwd")
        return da

    def corr_sums(self, x, y):
        self._need(x, y)
        acc = torch.zeros(5, dtype=torch.float64, device=x.device)
        _lib.check(self.L.mra_corr_sums(_ptr(x), _ptr(y), x.numel(), _dt(x), _ptr(acc), self._stream()),
                   "mra_corr_sums")
        return acc

    # -- optimiser --------------------------------------------------------------------------
    def adam_step(self, params, grads, exp_avgs, exp_avg_sqs, shadows, lr, beta1, beta2, eps, step):
        arr = (_lib.AdamTensor * len(params))()
        for i, (p, g, m, v, s) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs, shadows)):
            self._need_dense(p, g, m, v, s)
            arr[i].p, arr[i].g, arr[i].m, arr[i].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
            arr[i].shadow = s.data_ptr() if s is not None else None
            arr[i].numel = p.numel()
        _lib.check(self.L.mra_adam_multi(arr, len(params), lr, beta1, beta2, eps, step, self._stream()),
                   "mra_adam_multi")

    def adam_step_dev(self, params, grads, exp_avgs, exp_avg_sqs, shadows, hyper):
        """Same sweep, hyper-parameters {lr, b1, b2, eps, lr/(1-b1^t), sqrt(1-b2^t)} read from the device tensor ``hyper``."""
        arr = (_lib.AdamTensor * len(params))()
        for i, (p, g, m, v, s) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs, shadows)):
            self._need_dense(p, g, m, v, s)
            arr[i].p, arr[i].g, arr[i].m, arr[i].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
            arr[i].shadow = s.data_ptr() if s is not None else None
            arr[i].numel = p.numel()
        assert hyper.is_cuda and hyper.dtype == torch.float32 and hyper.numel() >= 6
        _lib.check(self.L.mra_adam_multi_dev(arr, len(params), _ptr(hyper), self._stream()), "mra_adam_multi_dev")

    def adam_advance(self, state, hyper):
        """t += 1 on the device and hyper <- {lr, b1, b2, eps, lr/(1-b1^t), sqrt(1-b2^t)} (see mra_adam_advance)."""
        assert state.is_cuda and state.dtype == torch.float64 and state.numel() >= 5
        assert hyper.is_cuda and hyper.dtype == torch.float32 and hyper.numel() >= 6
        _lib.check(self.L.mra_adam_advance(_ptr(state), _ptr(hyper), self._stream()), "mra_adam_advance")

    @staticmethod
    def _need_dense(p, g, m, v, s):
        for t in (p, g, m, v, s):
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("mra_gan_b200 Adam runs on CUDA tensors only (no CPU fallback)")
            if t.stride() != p.stride() or t.shape != p.shape:
                raise RuntimeError("Adam state / grad must share the parameter's memory layout")
        if p.dtype != torch.float32 or g.dtype != torch.float32:
            raise RuntimeError("Adam master parameters and grads are fp32")

    # -- sliding-window helpers -------------------------------------------------------------
    def window_extract(self, vol, i0, j0, k0, patch, dtype):
        self._need(vol)
        X, Y, Z = vol.shape
        out = torch.empty((1,) + tuple(patch) + (1,), dtype=dtype, device=vol.device)
        _lib.check(self.L.mra_window_extract(_ptr(vol), X, Y, Z, i0, j0, k0, patch[0], patch[1], patch[2], _ptr(out),
                                             _dt(out), self._stream()), "mra_window_extract")
        return out

    def window_accumulate(self, pred, label, weight, i0, j0, k0):
        self._need(pred, label, weight)
        X, Y, Z = label.shape
        px, py, pz = pred.shape[1:4]
        _lib.check(self.L.mra_window_accumulate(_ptr(pred), _dt(pred), _ptr(label), _ptr(weight), X, Y, Z, i0, j0, k0,
                                                px, py, pz, self._stream()), "mra_window_accumulate")

    def window_finalize(self, label, weight):
        self._need(label, weight)
        _lib.check(self.L.mra_window_finalize(_ptr(label), _ptr(weight), label.numel(), self._stream()),
                   "mra_window_finalize")

    def tc_error(self, reset=True):
        return int(self.L.mra_debug_tc_error(int(reset)))

    def debug_counters(self, reset=True):
        buf = (C.c_ulonglong * 8)()
        _lib.check(self.L.mra_debug_counters(buf, int(reset)), "mra_debug_counters")
        return list(buf)

    def launch_count(self):
        return int(self.L.mra_debug_launch_count())


_impl = None


def impl():
    """The active implementation.  Created lazily; needs the built library AND a CUDA device."""
    global _impl
    if _impl is None:
        if not torch.cuda.is_available():
            raise RuntimeError("mra_gan_b200 needs a CUDA device (sm_100a): there is no CPU fallback")
        _impl = CudaImpl()
    return _impl


def set_impl(obj):
    """Test hook: install another implementation object (the CPU test-suite's emulator)."""
    global _impl
    prev, _impl = _impl, obj
    return prev


def plan_describe(g, n, in_dims, which, dtype=MRA_BF16):
    """Host-only: the implicit-GEMM plan of a conv as int32 words (no GPU needed)."""
    L = _lib.lib()
    d = _lib.ConvDesc()
    d.n, d.cin, d.cout = n, g.cin, g.cout
    d.din, d.hin, d.win = in_dims
    d.dout, d.hout, d.wout = g.out_dims(tuple(in_dims))
    d.k, d.stride, d.pad, d.transposed, d.dtype = g.k, g.stride, g.pad, int(g.transposed), dtype
    buf = (C.c_int32 * 8192)()
    nwords = L.mra_conv_plan_describe(C.byref(d), which, buf, 8192)
    if nwords < 0:
        _lib.check(nwords, "mra_conv_plan_describe")
    return list(buf[:nwords])


def schedule_describe(g, n, in_dims, which, units=0, single=False, dtype=MRA_BF16, cap=1 << 22):
    """Host-only: the persistent schedule of the tensor-core kernel(s) of a conv op as int32 words, walked on the CPU by
    the functions the kernels run (mra_debug_schedule; layout in include/mra_gan_b200.h)."""
    L = _lib.lib()
    d = _lib.ConvDesc()
    d.n, d.cin, d.cout = n, g.cin, g.cout
    d.din, d.hin, d.win = in_dims
    d.dout, d.hout, d.wout = g.out_dims(tuple(in_dims))
    d.k, d.stride, d.pad, d.transposed, d.dtype = g.k, g.stride, g.pad, int(g.transposed), dtype
    buf = (C.c_int32 * cap)()
    nwords = L.mra_debug_schedule(C.byref(d), which, units, int(single), buf, cap)
    if nwords < 0:
        _lib.check(nwords, "mra_debug_schedule")
    return list(buf[:nwords])
