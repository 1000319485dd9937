"""GPU: achieved HBM bandwidth of the fused InstanceNorm kernels on the BASELINE shapes (bf16).
Each op is captured K times into a CUDA graph and replayed, so the figures are device time per call with no
host launch overhead (what the graph-replayed training step sees); the tensors of the small shapes stay in L2
between calls exactly as they do behind the producing conv.
Algorithmic bytes: fwd = read x + write y(+halo) [+ read residual]; bwd = read gy + read x + write dx [+ write dres]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mra_gan_b200 import ops
from mra_gan_b200.ops import ACT_RELU, ACT_NONE, ACT_LRELU
I = ops.impl()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = 10
CASES = [("G.c1 norm 64ch 128^3 pad0", 128, 64, 0, False), ("G.u2 norm 64ch 128^3 pad3", 128, 64, 3, False),
         ("G.d1 norm 128ch 64^3 pad0", 64, 128, 0, False), ("G.rb norm1 256ch 32^3 pad1", 32, 256, 1, False),
         ("G.rb norm2 256ch 32^3 pad1 +res", 32, 256, 1, True), ("D.2 norm 128ch 32^3", 32, 128, 0, False),
         ("D.3 norm 256ch 16^3", 16, 256, 0, False), ("D.4 norm 512ch 15^3", 15, 512, 0, False)]
def graph_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    return best
print("stream kernels: %s, batch %d" % (os.environ.get("MRA_NORM_STREAM", "1"), N))
for name, S, Cc, pad, res in CASES:
    x = torch.randn((N, S, S, S, Cc), device="cuda").to(torch.bfloat16)
    stats = I.inorm_stats(x)
    r = torch.randn((N, S + 2, S + 2, S + 2, Cc), device="cuda").to(torch.bfloat16) if res else None
    act = ACT_NONE if res else ACT_RELU
    y, mean, rstd = I.inorm_fwd(x, stats, r, pad, act, 0.0, 1 if res else -1)
    gy = torch.randn_like(y)
    e = x.numel() * 2; ep = y.numel() * 2
    t_f = graph_time(lambda: I.inorm_fwd(x, stats, r, pad, act, 0.0, 1 if res else -1))
    t_b = graph_time(lambda: I.inorm_bwd(gy, x, mean, rstd, pad, act, 0.0, 1 if res else -1))
    bf = e + ep + (r.numel() * 2 if res else 0)
    bb = ep + e + e + (r.numel() * 2 if res else 0)
    print("%-34s fwd %7.1f us %6.0f GB/s | bwd(stats+apply) %7.1f us %6.0f GB/s" % (name, t_f * 1e3, bf / t_f / 1e6, t_b * 1e3, bb / t_b / 1e6), flush=True)
    del x, y, gy, r
    torch.cuda.empty_cache()
print("tc/stream error flag:", I.tc_error())
