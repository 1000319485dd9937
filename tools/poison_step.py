"""GPU: uninitialised-read detector without compute-sanitizer.  Every torch.empty / empty_like on the device is
filled with 0xFF bytes (NaN in bf16 / fp32 / fp64) before use; a kernel that consumes memory nobody wrote turns the
losses into NaN.  Runs the tiny model (CUDA-core convs) and an ngf=64 model (tcgen05 convs + lowerings)."""
import os, random, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
_empty, _empty_like = torch.empty, torch.empty_like
def _poison(t):
    if t.is_cuda and t.numel():
        if t.is_contiguous():
            t.view(torch.uint8).fill_(0xFF)
        elif t.is_floating_point():
            t.fill_(float("nan"))
    return t
POISON = os.environ.get("POISON", "1") == "1"
if POISON:
    torch.empty = lambda *a, **k: _poison(_empty(*a, **k))
    torch.empty_like = lambda *a, **k: _poison(_empty_like(*a, **k))
from mra_gan_b200 import networks3D as N3
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle.ref_import import make_opt
for dt, ngf, size in ((torch.bfloat16, 8, 32), (torch.float32, 8, 32), (torch.bfloat16, 64, 32)):
    N3.set_default_compute_dtype(dt)
    opt = make_opt(ngf=ngf, ndf=ngf, pool_size=2, checkpoints_dir="/tmp/gc")
    random.seed(77); torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        m = create_model(opt); m.setup(opt)
    for step in range(2):
        A, B = OF.synthetic_patches(1, size, seed=200 + step)
        m.set_input([A, B]); m.optimize_parameters()
        print(str(dt), ngf, step, {k: round(v, 6) for k, v in m.get_current_losses().items()}, flush=True)
    with torch.no_grad():
        y = m.netG_A(A.cuda())
    print("   out finite:", bool(torch.isfinite(y).all()), "checksum %.6f" % float(y.double().abs().sum()), flush=True)
    del m
    torch.cuda.empty_cache()
