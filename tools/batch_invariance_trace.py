"""GPU: first instruction of the fused generator program whose output for sample b of a batch-B run differs from the
batch-1 run of that sample.  Usage: python tools/batch_invariance_trace.py [B] [b] [size]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import networks3D as N3, ops, functional as MF
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
N3.set_default_compute_dtype(torch.bfloat16)
torch.manual_seed(3)
g = N3.define_G(1, 1, 64, "resnet_9blocks", "instance").cuda()
xs = torch.rand(5, 1, S, S, S, device="cuda") * 2 - 1
prog = g.program()

def run(x):
    outs = []
    stats = None; saved = None; saved_pad = 0
    x = MF.to_channels_last(x, torch.bfloat16)
    for i, ins in enumerate(prog):
        op = ins[0]
        if op == "conv":
            _, m, act, slope, want_stats = ins
            x, stats = MF.ConvFn.apply(x, m.weight, m.bias, m, act, slope, want_stats, not want_stats, None)
            outs.append((i, "conv %d->%d k%d s%d" % (m.geom.cin, m.geom.cout, m.geom.k, m.geom.stride), x, stats))
        elif op == "norm":
            _, m, act, slope, pad, use_res = ins
            x = N3.apply_norm(x, stats, saved if use_res else None, m, act, slope, pad, saved_pad, None)
            stats = None
            outs.append((i, "norm act%d pad%d res%d" % (act, pad, use_res), x, None))
        elif op == "pad":
            x = MF.RepPadFn.apply(x, ins[1]); outs.append((i, "pad", x, None))
        elif op == "save":
            saved, saved_pad = x, ins[1]
        elif op == "act":
            x = MF.ActFn.apply(x, ins[1], ins[2]); outs.append((i, "act", x, None))
    return outs

with torch.no_grad():
    full = run(xs[:B])
    one = run(xs[b:b + 1])
for (i, name, xf, sf), (_, _, xo, so) in zip(full, one):
    e = float((xf[b].float() - xo[0].float()).abs().max())
    es = float((sf[b] - so[0]).abs().max()) if sf is not None else 0.0
    flag = " <== first difference" if (e > 0 or es > 0) else ""
    print("%3d %-28s out max-abs %.3e  stats max-abs %.3e%s" % (i, name, e, es, flag))
    if flag:
        if sf is not None:
            d = (sf[b] - so[0]).abs()
            print("    stats differ at", int((d > 0).sum()), "of", d.numel(), "entries; rel max", float((d / so[0].abs().clamp_min(1e-30)).max()))
        d = (xf[b].float() - xo[0].float()).abs()
        nz = (d > 0).nonzero()
        print("    differing elements:", nz.shape[0], "of", d.numel(), "first", nz[:3].tolist(), "last", nz[-3:].tolist())
        break
