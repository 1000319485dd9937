"""GPU (1..N ranks under torchrun): BASELINE config 5 -- sliding-window GA inference of a resnet_9blocks generator on a
synthetic 256x256x160 volume, 128^3 windows, windows sharded round-robin over the ranks, one reduce to rank 0.
Prints volume voxels/s and window voxels/s (device time, max over ranks) and checks the sharded result against the
single-rank result on rank 0."""
import contextlib, io, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mra_gan_b200 import networks3D as N3, parallel
from mra_gan_b200.models import create_model
from mra_gan_b200.inference import sliding_window_inference, window_grid

rank, world = parallel.init_distributed()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
N3.set_default_compute_dtype(torch.bfloat16)
ck = "/tmp/mra_cfg5_%d" % rank
os.makedirs(ck + "/cfg5", exist_ok=True)
opt = bench.make_opt(isTrain=False, model="test", model_suffix="", checkpoints_dir=ck, name="cfg5")
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(7)                                  # same weights on every rank
    g = N3.define_G(1, 1, 64, "resnet_9blocks", "instance")
    torch.save({k: v.detach().cpu().contiguous() for k, v in g.state_dict().items()}, ck + "/cfg5/latest_net_G.pth")
    tm = create_model(opt); tm.setup(opt)
vol = torch.rand(256, 256, 160, generator=torch.Generator().manual_seed(1234)) * 255
for stride in (64, 32):
    nwin = len(window_grid((256, 256, 160), (128, 128, 128), stride, stride))
    fn = lambda: sliding_window_inference(tm, vol, (128, 128, 128), stride, stride, rank, world, dtype=torch.float32)
    out = fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = sliding_window_inference(tm, vol, (128, 128, 128), stride, stride, 0, 1, dtype=torch.float32) if world > 1 else out
        err = float((out - ref).abs().max())
        t = float(ms) * 1e-3
        print("config 5, %d GPU(s), stride %d: %d windows, %.1f ms = %.1f Mvox/s (volume), %.1f Mvox/s (windows); "
              "max |sharded - single| = %.3g on the 0..255 scale" % (world, stride, nwin, t * 1e3, 256 * 256 * 160 / t / 1e6,
                                                                   nwin * 128 ** 3 / t / 1e6, err), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
