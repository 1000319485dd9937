"""GPU: achieved HBM bandwidth of the fused InstanceNorm kernels on the BASELINE shapes (batch 2, bf16).
Algorithmic bytes: fwd = read x + write y(+halo) [+ read residual]; bwd = read gy + read x + write dx [+ write dres]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ACT_RELU, ACT_NONE
I = ops.impl()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
CASES = [("G.c1 norm 64ch 128^3 pad0", 128, 64, 0, False), ("G.u2 norm 64ch 128^3 pad3", 128, 64, 3, False),
         ("G.d1 norm 128ch 64^3 pad0", 64, 128, 0, False), ("G.rb norm1 256ch 32^3 pad1", 32, 256, 1, False),
         ("G.rb norm2 256ch 32^3 pad1 +res", 32, 256, 1, True), ("D.2 norm 128ch 32^3", 32, 128, 0, False)]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps
for name, S, Cc, pad, res in CASES:
    x = torch.randn((N, S, S, S, Cc), device="cuda").to(torch.bfloat16)
    stats = I.inorm_stats(x)
    r = torch.randn((N, S + 2, S + 2, S + 2, Cc), device="cuda").to(torch.bfloat16) if res else None
    act = ACT_NONE if res else ACT_RELU
    y, mean, rstd = I.inorm_fwd(x, stats, r, pad, act, 0.0, 1 if res else -1)
    gy = torch.randn_like(y)
    e = x.numel() * 2; ep = y.numel() * 2
    t_f = timeit(lambda: I.inorm_fwd(x, stats, r, pad, act, 0.0, 1 if res else -1))
    t_b = timeit(lambda: I.inorm_bwd(gy, x, mean, rstd, pad, act, 0.0, 1 if res else -1))
    bf = e + ep + (r.numel() * 2 if res else 0)
    bb = ep + e + e + (r.numel() * 2 if res else 0)
    print("%-34s fwd %.3f ms %6.0f GB/s | bwd(stats+apply) %.3f ms %6.0f GB/s" % (name, t_f, bf / t_f / 1e6, t_b, bb / t_b / 1e6), flush=True)
    del x, y, gy, r
    torch.cuda.empty_cache()
