"""Sliding-window inference host logic (window grid, sharding, accumulation) against the fixture
produced by the reference's own test.py lines; ops are the CPU oracle implementation."""
import os

import numpy as np
import pytest
import torch

from mra_gan_b200 import inference
from mra_gan_b200 import networks3D as N3
from mra_gan_b200 import ops
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle import sliding_window as SW
from oracle.ops_ref import RefImpl
from oracle.ref_import import make_opt


@pytest.fixture(autouse=True)
def oracle_ops():
    prev = ops.set_impl(RefImpl(torch.float32))
    N3.set_default_compute_dtype(torch.float32)
    yield
    ops.set_impl(prev)
    N3.set_default_compute_dtype(torch.bfloat16)


def test_window_grid_equals_reference_loop():
    for shape, patch, s1, s2 in (((256, 256, 160), (128,) * 3, 32, 32), ((256, 256, 160), (128,) * 3, 64, 64),
                                 ((72, 64, 42), (32,) * 3, 16, 16), ((33, 40, 36), (32,) * 3, 7, 5)):
        ref = [(a, c, e) for (a, b, c, d, e, f) in SW.window_grid(shape, patch, s1, s2)]
        assert inference.window_grid(shape, patch, s1, s2) == ref
    assert len(inference.window_grid((256, 256, 160), (128,) * 3, 32, 32)) == 50


def _test_model(tmp_path, sd):
    ck = os.path.join(str(tmp_path), "sw")
    os.makedirs(ck, exist_ok=True)
    torch.save(sd, os.path.join(ck, "latest_net_G.pth"))
    opt = make_opt(ngf=8, isTrain=False, model="test", model_suffix="", checkpoints_dir=str(tmp_path), name="sw")
    m = create_model(opt)
    m.setup(opt)
    return m


def test_sliding_window_matches_reference_fixture(golden_dir, tmp_path):
    r = torch.load(os.path.join(golden_dir, "sliding_window_small.pt"), weights_only=False)
    sd = OF.make_weights(OF.resnet_g_spec(1, 1, 8, 9), r["weight_seed"], scale=r["weight_scale"])
    model = _test_model(tmp_path, sd)
    vol = np.random.RandomState(r["vol_seed"]).uniform(0, 255, size=r["shape"]).astype(np.float32)
    out = inference.sliding_window_inference(model, torch.from_numpy(vol), r["patch"], *r["stride"])
    assert tuple(out.shape) == tuple(r["shape"])
    # 0..255 scale.  5e-3 = 4e-5 on the generator's [-1, 1] output: the default loop sends two windows per pass, and the
    # CPU oracle's batched fp32 convolution sums in another order than its batch-1 call (2.1e-3 measured here)
    assert float((out - r["label"]).abs().max()) < 5e-3
    # sharding: the union of the ranks' partial sums equals the single-rank result
    parts = []
    for rank in range(3):
        lab = inference.sliding_window_inference.__wrapped__(model, torch.from_numpy(vol), r["patch"], *r["stride"],
                                                             rank=rank, world=3, _local_only=True)
        parts.append(lab)
    label = sum(p[0] for p in parts)
    weight = sum(p[1] for p in parts)
    merged = (label / weight + 0.01)[:, :, :r["shape"][2]]
    assert float((merged - r["label"]).abs().max()) < 5e-3


def test_windows_per_pass_does_not_change_the_result(golden_dir, tmp_path):
    """Several windows through the generator as one batch (per-sample instance statistics) == one window at a time."""
    r = torch.load(os.path.join(golden_dir, "sliding_window_small.pt"), weights_only=False)
    sd = OF.make_weights(OF.resnet_g_spec(1, 1, 8, 9), r["weight_seed"], scale=r["weight_scale"])
    model = _test_model(tmp_path, sd)
    vol = torch.from_numpy(np.random.RandomState(r["vol_seed"]).uniform(0, 255, size=r["shape"]).astype(np.float32))
    one = inference.sliding_window_inference(model, vol, r["patch"], *r["stride"], windows_per_pass=1)
    for wpp in (2, 3):
        many = inference.sliding_window_inference(model, vol, r["patch"], *r["stride"], windows_per_pass=wpp)
        # 0..255 scale; the CPU oracle's fp32 convolution sums in a batch-dependent order (1.6e-3 measured), the CUDA
        # kernels do not (tests/test_options_gpu.py holds the tight version of this check)
        assert float((many - one).abs().max()) < 5e-3
    assert float((one - r["label"]).abs().max()) < 2e-3


# ------------------------------------------------------------------------------------------------
# two real ranks (gloo): the windows of the grid are dealt out round-robin, every rank accumulates its own label /
# weight sums and ONE reduce(sum) to rank 0 finishes the volume (config 5's N > 1 path)
# ------------------------------------------------------------------------------------------------
def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _infer_worker(rank, world, port, tmp, golden, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from mra_gan_b200 import parallel
    assert parallel.init_distributed("gloo") == (rank, world)
    ops.set_impl(RefImpl(torch.float32))
    N3.set_default_compute_dtype(torch.float32)
    r = torch.load(os.path.join(golden, "sliding_window_small.pt"), weights_only=False)
    sd = OF.make_weights(OF.resnet_g_spec(1, 1, 8, 9), r["weight_seed"], scale=r["weight_scale"])
    model = _test_model(os.path.join(tmp, "rank%d" % rank), sd)
    vol = torch.from_numpy(np.random.RandomState(r["vol_seed"]).uniform(0, 255, size=r["shape"]).astype(np.float32))
    out = inference.sliding_window_inference(model, vol, r["patch"], *r["stride"], rank=rank, world=world)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sliding_window_reduces_to_the_reference_volume(golden_dir, tmp_path):
    import torch.multiprocessing as mp
    r = torch.load(os.path.join(golden_dir, "sliding_window_small.pt"), weights_only=False)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_infer_worker, args=(k, world, port, str(tmp_path), golden_dir, q)) for k in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out0 = torch.from_numpy(res[0])
    assert tuple(out0.shape) == tuple(r["shape"])                   # the odd-z pad is cut off again on every rank
    assert float((out0 - r["label"]).abs().max()) < 5e-3            # rank 0: the finished volume (0..255 scale)
    # the other rank keeps its partial label sums (not finalised): they must differ from the result -- i.e. the reduce
    # really went to rank 0 only and rank 1 did not silently compute the whole grid
    assert float((torch.from_numpy(res[1]) - r["label"]).abs().max()) > 1.0
