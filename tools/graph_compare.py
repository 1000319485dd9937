import os, random, sys, contextlib, io
sys.path.insert(0, "/root/repo")
import torch
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle.ref_import import make_opt
N3.set_default_compute_dtype(torch.bfloat16)
def run(use_graphs):
    opt = make_opt(ngf=8, ndf=8, pool_size=2, checkpoints_dir="/tmp/gc")
    random.seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        m = create_model(opt); m.setup(opt)
    for net, sd in zip((m.netG_A, m.netG_B, m.netD_A, m.netD_B), OF.build_cyclegan_weights(8, 8, seed=3)):
        net.load_state_dict({k: v.clone() for k, v in sd.items()})
    if use_graphs: m.enable_cuda_graphs(warmup_steps=1)
    out = []
    for step in range(6):
        A, B = OF.synthetic_patches(1, 32, seed=200 + step)
        m.set_input([A, B]); m.optimize_parameters()
        l = m.get_current_losses()
        out.append([l[k] for k in ("D_A", "G_A", "cycle_A", "idt_A", "D_B", "G_B")])
    return out
a, b, c = run(False), run(False), run(True)
for s in range(6):
    print("step", s)
    print("  eager  ", ["%.5f" % v for v in a[s]])
    print("  eager2 ", ["%.5f" % v for v in b[s]])
    print("  graphs ", ["%.5f" % v for v in c[s]])
