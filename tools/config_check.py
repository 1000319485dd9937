"""GPU: BASELINE configs 4 and 5 at full size (timing + sanity), single GPU.
  config 4: unet_128 (7-down UNet) CycleGAN step on 128^3 patches, bf16
  config 5: sliding-window inference of a resnet_9blocks generator on a synthetic 256x256x160 volume, 128^3 windows"""
import contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mra_gan_b200 import networks3D as N3, ops
from mra_gan_b200.models import create_model
from mra_gan_b200.inference import sliding_window_inference, window_grid
N3.set_default_compute_dtype(torch.bfloat16)
I = ops.impl()

def timed(fn, reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

# ---- config 4
batch = 2
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(bench.make_opt(netG="unet_128")); m.setup(bench.make_opt(netG="unet_128"))
A = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
B = torch.rand(batch, 1, 128, 128, 128, device="cuda") * 2 - 1
def step():
    m.set_input([A, B]); m.optimize_parameters()
for _ in range(3): step()
t = timed(step, 3)
print("config 4: unet_128 CycleGAN step, batch %d, 128^3: %.1f ms/step = %.2f Mvox/s ; losses %s ; tc flag %d" % (
    batch, t * 1e3, batch * 128 ** 3 / t / 1e6, {k: round(v, 3) for k, v in m.get_current_losses().items()}, I.tc_error()), flush=True)
del m, A, B
torch.cuda.empty_cache()

# ---- config 5
opt = bench.make_opt(isTrain=False, model="test", model_suffix="", checkpoints_dir="/tmp/mra_cfg5", name="cfg5")
os.makedirs("/tmp/mra_cfg5/cfg5", exist_ok=True)
with contextlib.redirect_stdout(io.StringIO()):
    g = N3.define_G(1, 1, 64, "resnet_9blocks", "instance")
    torch.save({k: v.detach().cpu().contiguous() for k, v in g.state_dict().items()}, "/tmp/mra_cfg5/cfg5/latest_net_G.pth")
    tm = create_model(opt); tm.setup(opt)
gen = torch.Generator().manual_seed(1234)
vol = torch.rand(256, 256, 160, generator=gen) * 255
for stride in (64, 32):
    nwin = len(window_grid((256, 256, 160), (128, 128, 128), stride, stride))
    fn = lambda: sliding_window_inference(tm, vol, (128, 128, 128), stride, stride, dtype=torch.float32)
    out = fn()
    t = timed(fn, 1)
    print("config 5: sliding window 256x256x160, 128^3 windows, stride %d: %d windows, %.1f ms = %.2f Mvox/s (volume), %.2f Mvox/s (windows); "
          "out range [%.1f, %.1f] finite %s ; tc flag %d" % (stride, nwin, t * 1e3, 256 * 256 * 160 / t / 1e6, nwin * 128 ** 3 / t / 1e6,
                                                             float(out.min()), float(out.max()), bool(torch.isfinite(out).all()), I.tc_error()), flush=True)
