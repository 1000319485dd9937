"""GPU: is a sample's result independent of its batch mates?  The BASELINE generator (ngf = 64, bf16, every conv on the
tensor-core path) is run on B different windows at once and each sample is compared with its own batch-1 run; then the
same per layer type for the convolution ops (fprop / dgrad / wgrad) at the BASELINE shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import networks3D as N3, ops
from mra_gan_b200.ops import ConvGeom
I = ops.impl()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N3.set_default_compute_dtype(torch.bfloat16)
torch.manual_seed(3)
g = N3.define_G(1, 1, 64, "resnet_9blocks", "instance").cuda()
xs = torch.rand(5, 1, S, S, S, device="cuda") * 2 - 1
with torch.no_grad():
    ones = [g(xs[i:i + 1]) for i in range(5)]
    for B in (2, 3, 4, 5):
        yb = g(xs[:B])
        errs = [float((yb[i:i + 1] - ones[i]).abs().max()) for i in range(B)]
        print("generator %d^3 batch %d: per-sample max-abs vs batch-1 run:" % (S, B), ["%.2e" % e for e in errs], flush=True)
print("tc error", I.tc_error())
if os.environ.get("LAYERS", "1") == "1":
    LAYERS = [("G.c1", ConvGeom(1, 64, 7, 1, 0), S + 6), ("G.d1", ConvGeom(64, 128, 3, 2, 1), S), ("G.d2", ConvGeom(128, 256, 3, 2, 1), S // 2),
              ("G.rb", ConvGeom(256, 256, 3, 1, 0), S // 4 + 2), ("G.u1", ConvGeom(256, 128, 3, 2, 1, True, 1), S // 4),
              ("G.u2", ConvGeom(128, 64, 3, 2, 1, True, 1), S // 2), ("G.c4", ConvGeom(64, 1, 7, 1, 0), S + 6)]
    for name, geo, d in LAYERS:
        dims = (d,) * 3
        od = geo.out_dims(dims)
        x = torch.randn((5,) + dims + (geo.cin,), device="cuda").to(torch.bfloat16)
        w = (torch.randn((geo.taps, geo.cout, geo.cin), device="cuda") * 0.05).to(torch.bfloat16)
        wT = I.pack_weight_t(w, torch.bfloat16)
        dy = torch.randn((5,) + od + (geo.cout,), device="cuda").to(torch.bfloat16)
        f1 = [I.conv_fprop(x[i:i + 1].contiguous(), w, None, geo, want_stats=geo.cout > 1) for i in range(5)]
        d1 = [I.conv_dgrad(dy[i:i + 1].contiguous(), wT, geo, dims) for i in range(5)]
        for B in (2, 3, 4, 5):
            yb, sb = I.conv_fprop(x[:B].contiguous(), w, None, geo, want_stats=geo.cout > 1)
            db = I.conv_dgrad(dy[:B].contiguous(), wT, geo, dims)
            ef = max(float((yb[i].float() - f1[i][0][0].float()).abs().max()) for i in range(B))
            es = max(float((sb[i] - f1[i][1][0]).abs().max() / (f1[i][1][0].abs().max() + 1e-30)) for i in range(B)) if sb is not None else 0.0
            ed = max(float((db[i].float() - d1[i][0].float()).abs().max()) for i in range(B))
            dwb, _ = I.conv_wgrad(x[:B].contiguous(), dy[:B].contiguous(), geo)
            dws = sum(I.conv_wgrad(x[i:i + 1].contiguous(), dy[i:i + 1].contiguous(), geo)[0] for i in range(B))
            ew = float((dwb - dws).abs().max() / (dws.abs().max() + 1e-30))
            print("%-5s batch %d: fprop %.2e stats(rel) %.2e dgrad %.2e wgrad(rel to sum of batch-1) %.2e" % (name, B, ef, es, ed, ew), flush=True)
        del x, dy, f1, d1
        torch.cuda.empty_cache()
    print("tc error", I.tc_error())
