"""Diagnostic (GPU): run the tcgen05 conv kernels against the CUDA-core kernels on the same bf16
inputs and print WHERE they differ (per op, per channel block, per row-in-tile), so that one gpurun
call is enough to localise a descriptor / swizzle / indexing bug.  Writes gpurun_out/tc_diag.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops  # noqa: E402
from mra_gan_b200.ops import ConvGeom  # noqa: E402

CASES = [
    (ConvGeom(64, 64, 1, 1, 0), (4, 4, 8), 1),
    (ConvGeom(64, 64, 3, 1, 0), (6, 6, 10), 1),
    (ConvGeom(128, 256, 3, 1, 0), (10, 10, 10), 1),
    (ConvGeom(256, 256, 3, 1, 0), (10, 10, 10), 2),
    (ConvGeom(64, 128, 3, 2, 1), (16, 16, 16), 1),
    (ConvGeom(64, 128, 4, 2, 1), (16, 16, 16), 2),
    (ConvGeom(128, 512, 4, 1, 1), (10, 10, 10), 1),
    (ConvGeom(256, 128, 3, 2, 1, True, 1), (6, 6, 6), 2),
    (ConvGeom(128, 64, 4, 2, 1, True, 0), (5, 5, 5), 1),
    (ConvGeom(1, 64, 7, 1, 0), (14, 15, 17), 2),
    (ConvGeom(64, 1, 7, 1, 0), (14, 15, 17), 2),
]


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def describe(name, got, want, out):
    r = rel(got, want)
    out.append("  %-6s rel_l2=%.3e  max|got|=%.3e max|want|=%.3e nan=%d" % (
        name, r, float(got.abs().max()), float(want.abs().max()), int(torch.isnan(got.float()).sum())))
    if r > 2e-2:
        g, w = got.double(), want.double()
        if g.dim() == 5:                       # activation (N,D,H,W,C)
            err = (g - w).abs()
            out.append("    err by channel block of 32: " + " ".join(
                "%.2e" % float(err[..., c:c + 32].mean()) for c in range(0, g.shape[-1], 32)))
            flat = err.mean(-1).flatten()
            out.append("    err by position (first 64): " + " ".join("%.1e" % float(v) for v in flat[:64]))
            ratio = (g.flatten()[:16] / (w.flatten()[:16] + 1e-30))
            out.append("    got/want first 16: " + " ".join("%.3f" % float(v) for v in ratio))
        else:                                  # weights [taps][co][ci]
            err = (g - w).abs()
            out.append("    err by tap: " + " ".join("%.2e" % float(err[t].mean()) for t in range(g.shape[0])))
            out.append("    err by co block of 32: " + " ".join(
                "%.2e" % float(err[:, c:c + 32].mean()) for c in range(0, g.shape[1], 32)))
            out.append("    err by ci block of 32: " + " ".join(
                "%.2e" % float(err[:, :, c:c + 32].mean()) for c in range(0, g.shape[2], 32)))


def main():
    I = ops.impl()
    out = []
    for g, dims, n in CASES:
        gen = torch.Generator().manual_seed(0)
        x = torch.randn((n,) + dims + (g.cin,), generator=gen).to(torch.bfloat16).cuda()
        w = (torch.randn((g.taps, g.cout, g.cin), generator=gen) / (g.taps * g.cin) ** 0.5).to(torch.bfloat16).cuda()
        dy = torch.randn((n,) + g.out_dims(dims) + (g.cout,), generator=gen).to(torch.bfloat16).cuda()
        wT = I.pack_weight_t(w, torch.bfloat16)
        out.append("case cin=%d cout=%d k=%d s=%d p=%d T=%d dims=%s n=%d" % (
            g.cin, g.cout, g.k, g.stride, g.pad, g.transposed, dims, n))
        res = {}
        for naive in (True, False):
            I.force_naive = naive
            try:
                y, st = I.conv_fprop(x, w, None, g, want_stats=True)
                dx = I.conv_dgrad(dy, wT, g, dims)
                dw, _ = I.conv_wgrad(x, dy, g)
                torch.cuda.synchronize()
                res[naive] = (y, st, dx, dw)
            except Exception as e:  # noqa: BLE001
                out.append("  EXCEPTION (naive=%s): %r" % (naive, e))
            err = I.tc_error()
            if err:
                out.append("  tc error flag = %d (naive=%s)" % (err, naive))
        I.force_naive = False
        if True in res and False in res:
            for name, a, b in zip(("fprop", "stats", "dgrad", "wgrad"), res[False], res[True]):
                describe(name, a, b, out)
    txt = "\n".join(out)
    print(txt)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_diag.txt", "w") as f:
        f.write(txt + "\n")


if __name__ == "__main__":
    main()
