"""Op-level oracle (TEST INFRASTRUCTURE): the semantics of every C-ABI op in
include/mra_gan_b200.h restated with torch CPU functional calls -- the same ATen calls the
reference's nn.Modules make (models/networks3D.py), wrapped in the library's data conventions
(channels-last activations (N,D,H,W,C), packed weights [taps][Cout][Cin]).

``RefImpl`` has the interface of ``mra_gan_b200.ops.CudaImpl`` so that (a) the GPU tests can compare
op by op and (b) the CPU test-suite can install it (``ops.set_impl``) to exercise the package's
host logic (module wiring, autograd formulas, fusion plan, data-parallel buckets) without a GPU.
It is never reachable from the product path.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
LOSS_L1, LOSS_MSE_CONST, LOSS_BCE_CONST = 0, 1, 2


def to_ncdhw(t):
    return t.permute(0, 4, 1, 2, 3)


def to_ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def pack_weight(w_ref, transposed=False):
    """reference layout (Cout,Cin,k,k,k) [ConvTranspose3d: (Cin,Cout,k,k,k)] -> [taps][Cout][Cin]"""
    k = w_ref.shape[2]
    perm = (2, 3, 4, 1, 0) if transposed else (2, 3, 4, 0, 1)
    p = w_ref.permute(*perm).contiguous()
    return p.view(k ** 3, p.shape[3], p.shape[4])


def unpack_weight(w_packed, k, transposed=False):
    t, co, ci = w_packed.shape
    v = w_packed.view(k, k, k, co, ci)
    return v.permute(4, 3, 0, 1, 2) if transposed else v.permute(3, 4, 0, 1, 2)


def _act(v, act, slope):
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_LRELU:
        return F.leaky_relu(v, slope)
    if act == ACT_TANH:
        return torch.tanh(v)
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    return v


def _act_grad_from_output(y, act, slope):
    if act == ACT_RELU:
        return (y > 0).to(y.dtype)
    if act == ACT_LRELU:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    if act == ACT_TANH:
        return 1 - y * y
    if act == ACT_SIGMOID:
        return y * (1 - y)
    return torch.ones_like(y)


def _fold_pad(g, pad):
    """backward of replication padding on a channels-first tensor (N,C,Dp,Hp,Wp)."""
    if pad == 0:
        return g
    for dim in (2, 3, 4):
        n = g.shape[dim] - 2 * pad
        lo = g.narrow(dim, 0, pad + 1).sum(dim, keepdim=True)
        hi = g.narrow(dim, n + pad - 1, pad + 1).sum(dim, keepdim=True)
        if n == 1:
            g = g.sum(dim, keepdim=True)
        else:
            mid = g.narrow(dim, pad + 1, n - 2)
            g = torch.cat([lo, mid, hi], dim)
    return g


class RefImpl:
    name = "oracle"

    def __init__(self, compute_dtype=torch.float32):
        self.cd = compute_dtype

    def _c(self, t):
        return None if t is None else t.to(self.cd)

    # -- convolution family (F.conv3d / F.conv_transpose3d and their autograd) ---------------
    def _conv(self, xc, wp, bias, g):
        w = unpack_weight(wp, g.k, g.transposed)
        if g.transposed:
            return F.conv_transpose3d(xc, w, bias, stride=g.stride, padding=g.pad, output_padding=g.output_padding)
        return F.conv3d(xc, w, bias, stride=g.stride, padding=g.pad)

    def conv_fprop(self, x, w, bias, g, act=ACT_NONE, slope=0.2, want_stats=False):
        y = self._conv(to_ncdhw(self._c(x)), self._c(w), self._c(bias), g)
        stats = None
        if want_stats:
            yd = y.double()
            stats = torch.stack([yd.sum((2, 3, 4)), (yd * yd).sum((2, 3, 4))], -1)
        return to_ndhwc(_act(y, act, slope)).to(x.dtype), stats

    def _conv_backward(self, gy, x, w_ref, g, mask):
        """ATen's convolution_backward -- the node torch's autograd runs for nn.Conv3d / nn.ConvTranspose3d -- called
        directly (usable below the autograd dispatch key, e.g. from inside a torch.library operator)."""
        op = [g.output_padding] * 3 if g.transposed else [0] * 3
        return torch.ops.aten.convolution_backward(gy.contiguous(), x, w_ref.contiguous(), None, [g.stride] * 3, [g.pad] * 3,
                                                   [1] * 3, g.transposed, op, 1, mask)

    def conv_dgrad(self, dy, wT, g, in_dims):
        wp = self._c(wT).transpose(1, 2).contiguous()
        x = torch.zeros((dy.shape[0], g.cin) + tuple(in_dims), dtype=self.cd)
        dx = self._conv_backward(to_ncdhw(self._c(dy)), x, unpack_weight(wp, g.k, g.transposed), g, [True, False, False])[0]
        return to_ndhwc(dx).to(dy.dtype)

    def conv_dgrad_nstats_supported(self, g, n, in_dims, dtype):
        return True

    def conv_dgrad_nstats(self, dy, wT, g, in_dims, y_act, norm_act, norm_slope, ws=None, reuse=False):
        """mra_conv3d_dgrad_nstats: dx plus {sum dx * act'(y), sum dx * y} over ALL (padded) positions of y_act."""
        dx = self.conv_dgrad(dy, wT, g, in_dims)
        ns = {ACT_NONE: 1.0, ACT_RELU: 0.0}.get(norm_act, norm_slope)
        ya, gx = self._c(y_act), self._c(dx)
        s0 = (gx * torch.where(ya > 0, torch.ones_like(ya), torch.full_like(ya, ns))).sum((1, 2, 3))
        s1 = (gx * ya).sum((1, 2, 3))
        return dx, torch.stack([s0, s1], -1).double()

    def conv_wgrad(self, x, dy, g, want_bias=False, acc_dw=None, acc_db=None):
        wp = torch.zeros((g.taps, g.cout, g.cin), dtype=self.cd)
        gw = self._conv_backward(to_ncdhw(self._c(dy)), to_ncdhw(self._c(x)).contiguous(),
                                 unpack_weight(wp, g.k, g.transposed), g, [False, True, False])[1]
        dw = pack_weight(gw, g.transposed)
        db = self._c(dy).sum((0, 1, 2, 3)).float() if want_bias else None
        if acc_dw is not None:                       # MRA_CONV_ACCUMULATE: add into the caller's buffers
            acc_dw += dw.float()
            if want_bias:
                acc_db += db
            return acc_dw, (acc_db if want_bias else None)
        return dw.float().contiguous(), db

    def conv_uses_tensor_cores(self, g, n, in_dims, dtype, which):
        return False

    def pack_weight_t(self, w, dst_dtype):
        return w.transpose(1, 2).contiguous().to(dst_dtype)

    def convert(self, t, dst_dtype):
        return t.to(dst_dtype)

    # -- instance norm family ----------------------------------------------------------------
    def inorm_stats(self, x):
        xd = x.double()
        return torch.stack([xd.sum((1, 2, 3)), (xd * xd).sum((1, 2, 3))], -1)

    def inorm_fwd(self, x, stats, residual=None, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1, eps=1e-5,
                  momentum=0.1, running_mean=None, running_var=None, use_running=False, given=None):
        n, d, h, w, c = x.shape
        V = d * h * w
        if given is not None:
            mean, rstd = given[0].to(self.cd), given[1].to(self.cd)
        elif use_running:
            mean = running_mean.to(self.cd).expand(n, c).contiguous()
            rstd = (1.0 / torch.sqrt(running_var.to(self.cd) + eps)).expand(n, c).contiguous()
        else:
            mu = stats[..., 0] / V
            var = (stats[..., 1] / V - mu * mu).clamp_min(0)
            mean = mu.to(self.cd)
            rstd = 1.0 / torch.sqrt(var.to(self.cd) + eps)
            if running_mean is not None:
                running_mean.mul_(1 - momentum).add_(momentum * mu.mean(0).to(running_mean.dtype))
                running_var.mul_(1 - momentum).add_(momentum * (var * V / (V - 1)).mean(0).to(running_var.dtype))
        xc = self._c(x)
        y = _act((xc - mean.view(n, 1, 1, 1, c)) * rstd.view(n, 1, 1, 1, c), act, slope)
        if residual is not None:
            r = self._c(residual)
            if res_pad > 0:
                r = r[:, res_pad:-res_pad, res_pad:-res_pad, res_pad:-res_pad, :]
            y = y + r
        if pad > 0:
            y = to_ndhwc(F.pad(to_ncdhw(y), (pad,) * 6, mode="replicate"))
        return y.contiguous().to(x.dtype), mean.float(), rstd.float()

    def inorm_bwd(self, gy, x, mean, rstd, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1, use_running=False):
        n, d, h, w, c = x.shape
        g = to_ndhwc(_fold_pad(to_ncdhw(self._c(gy)), pad))
        m, r = mean.to(self.cd).view(n, 1, 1, 1, c), rstd.to(self.cd).view(n, 1, 1, 1, c)
        xh = (self._c(x) - m) * r
        dy = g * _act_grad_from_output(xh, act, slope)
        if use_running:
            dx = dy * r
        else:
            m1 = dy.mean((1, 2, 3), keepdim=True)
            m2 = (dy * xh).mean((1, 2, 3), keepdim=True)
            dx = r * (dy - m1 - xh * m2)
        dres = None
        if res_pad >= 0:
            dres = F.pad(g, (0, 0) + (res_pad,) * 6).contiguous().to(x.dtype)
        return dx.contiguous().to(x.dtype), dres

    def inorm_bwd_stats(self, gy, x, mean, rstd, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1):
        n, d, h, w, c = x.shape
        g = to_ndhwc(_fold_pad(to_ncdhw(self._c(gy)), pad))
        m, r = mean.to(self.cd).view(n, 1, 1, 1, c), rstd.to(self.cd).view(n, 1, 1, 1, c)
        xh = (self._c(x) - m) * r
        dy = g * _act_grad_from_output(xh, act, slope)
        return torch.stack([dy.sum((1, 2, 3)), (dy * xh).sum((1, 2, 3))], -1).double()

    def inorm_bwd_apply(self, gy, x, mean, rstd, sums, pad=0, act=ACT_NONE, slope=0.2, res_pad=-1):
        n, d, h, w, c = x.shape
        V = d * h * w
        g = to_ndhwc(_fold_pad(to_ncdhw(self._c(gy)), pad))
        m, r = mean.to(self.cd).view(n, 1, 1, 1, c), rstd.to(self.cd).view(n, 1, 1, 1, c)
        xh = (self._c(x) - m) * r
        dy = g * _act_grad_from_output(xh, act, slope)
        m1 = (sums[..., 0] / V).to(self.cd).view(n, 1, 1, 1, c)
        m2 = (sums[..., 1] / V).to(self.cd).view(n, 1, 1, 1, c)
        dx = r * (dy - m1 - xh * m2)
        dres = None
        if res_pad >= 0:
            dres = F.pad(g, (0, 0) + (res_pad,) * 6).contiguous().to(x.dtype)
        return dx.contiguous().to(x.dtype), dres

    def act_fwd(self, x, act, slope=0.2):
        return _act(self._c(x), act, slope).to(x.dtype)

    def act_bwd(self, dy, y, act, slope=0.2):
        return (self._c(dy) * _act_grad_from_output(self._c(y), act, slope)).to(dy.dtype)

    def cat2_act_fwd(self, a, b, act, slope=0.0):
        return _act(torch.cat([self._c(a), self._c(b)], 4), act, slope).contiguous().to(a.dtype)

    def cat2_act_bwd(self, dout, out, ca, act, slope=0.0, want=(True, True)):
        g = self._c(dout) * _act_grad_from_output(self._c(out), act, slope)
        da = g[..., :ca].contiguous().to(out.dtype) if want[0] else None
        db = g[..., ca:].contiguous().to(out.dtype) if want[1] else None
        return da, db

    def mask_scale(self, x, keep, scale):
        return (self._c(x) * keep.to(self.cd) * scale).to(x.dtype)

    def reppad_fwd(self, x, pad):
        return to_ndhwc(F.pad(to_ncdhw(x), (pad,) * 6, mode="replicate"))

    def reppad_bwd(self, gy, pad):
        return to_ndhwc(_fold_pad(to_ncdhw(self._c(gy)), pad)).to(gy.dtype)

    # -- losses ------------------------------------------------------------------------------
    def loss_fwd(self, kind, a, b=None, target=0.0):
        ac = self._c(a)
        if kind == LOSS_L1:
            return F.l1_loss(ac, self._c(b)).float()
        t = torch.full_like(ac, target)
        return (F.mse_loss(ac, t) if kind == LOSS_MSE_CONST else F.binary_cross_entropy(ac, t)).float()

    def loss_bwd(self, kind, a, b, target, gout, scale):
        ac = self._c(a)
        if kind == LOSS_L1:
            v = torch.sign(ac - self._c(b))
        elif kind == LOSS_MSE_CONST:
            v = 2 * (ac - target)
        else:
            v = (ac - target) / ((1 - ac) * ac).clamp_min(1e-12)
        return (v * (gout.to(self.cd) * scale)).to(a.dtype)

    def corr_sums(self, x, y):
        xd, yd = x.double().flatten(), y.double().flatten()
        return torch.stack([xd.sum(), yd.sum(), (xd * yd).sum(), (xd * xd).sum(), (yd * yd).sum()])

    # -- optimiser (torch.optim.Adam single-tensor arithmetic) -----------------------------------
    def adam_step(self, params, grads, exp_avgs, exp_avg_sqs, shadows, lr, beta1, beta2, eps, step):
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        for p, g, m, v, s in zip(params, grads, exp_avgs, exp_avg_sqs, shadows):
            m.lerp_(g, 1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
            p.addcdiv_(m, denom, value=-(lr / bc1))
            if s is not None:
                s.copy_(p)

    # -- sliding window ----------------------------------------------------------------------------
    def window_extract(self, vol, i0, j0, k0, patch, dtype):
        px, py, pz = patch
        w = (vol[i0:i0 + px, j0:j0 + py, k0:k0 + pz] - 127.5) / 127.5
        return w.reshape(1, px, py, pz, 1).to(dtype)

    def window_accumulate(self, pred, label, weight, i0, j0, k0):
        px, py, pz = pred.shape[1:4]
        label[i0:i0 + px, j0:j0 + py, k0:k0 + pz] += pred.float().reshape(px, py, pz) * 127.5 + 127.5
        weight[i0:i0 + px, j0:j0 + py, k0:k0 + pz] += 1.0

    def window_finalize(self, label, weight):
        label.copy_(label / weight + 0.01)

    def tc_error(self, reset=True):
        return 0
