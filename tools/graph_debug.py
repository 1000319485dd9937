import os, random, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle.ref_import import make_opt
N3.set_default_compute_dtype(torch.bfloat16)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
def run(use_graphs):
    opt = make_opt(ngf=8, ndf=8, pool_size=2, checkpoints_dir="/tmp/gc")
    random.seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        m = create_model(opt); m.setup(opt)
    for net, sd in zip((m.netG_A, m.netG_B, m.netD_A, m.netD_B), OF.build_cyclegan_weights(8, 8, seed=3)):
        net.load_state_dict({k: v.clone() for k, v in sd.items()})
    if use_graphs: m.enable_cuda_graphs(warmup_steps=1)
    out = []
    for step in range(STEPS):
        A, B = OF.synthetic_patches(1, 32, seed=200 + step)
        m.set_input([A, B]); m.optimize_parameters()
        l = m.get_current_losses()
        out.append([l[k] for k in ("D_A", "G_A", "cycle_A", "idt_A", "D_B", "G_B")])
    torch.cuda.synchronize()
    sd = {k: v.detach().float().cpu().clone() for k, v in m.netG_A.state_dict().items()}
    with torch.no_grad():
        y1 = m.netG_A(A.cuda()).float().cpu()
        y2 = m.netG_A(A.cuda()).float().cpu()
    return out, sd, y1, y2
(a, sa, ya1, ya2), (c, sc, yc1, yc2) = run(False), run(True)
for s in range(STEPS):
    print("step", s, "eager ", ["%.5f" % v for v in a[s]])
    print("step", s, "graphs", ["%.5f" % v for v in c[s]])
print("eval twice: eager %.3e graphs %.3e ; graphs vs eager: first %.4f second %.4f" % (
    OF.rel_l2(ya2, ya1), OF.rel_l2(yc2, yc1), OF.rel_l2(yc1, ya1), OF.rel_l2(yc2, ya2)))
worst = sorted(((OF.rel_l2(sc[k], sa[k]) if sa[k].numel() else 0.0, k) for k in sa), reverse=True)[:6]
print("state_dict diffs:", [(k, "%.3e" % v) for v, k in worst])
