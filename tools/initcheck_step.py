"""GPU: two eager CycleGAN steps of the tiny model (for `compute-sanitizer --tool initcheck / memcheck`)."""
import os, random, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle.ref_import import make_opt
N3.set_default_compute_dtype(torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32)
ngf = int(os.environ.get("NGF", "8"))
opt = make_opt(ngf=ngf, ndf=ngf, pool_size=2, checkpoints_dir="/tmp/gc")
random.seed(77)
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(opt); m.setup(opt)
for step in range(2):
    A, B = OF.synthetic_patches(1, 32, seed=200 + step)
    m.set_input([A, B]); m.optimize_parameters()
    print(step, {k: round(v, 5) for k, v in m.get_current_losses().items()}, flush=True)
