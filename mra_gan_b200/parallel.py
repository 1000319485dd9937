"""Patch-wise data parallelism: one process per GPU, bucketed gradient all-reduce overlapped with
the backward pass (the reference is single-device: nn.DataParallel is commented out,
models/networks3D.py:69-75; SURVEY.md 8e defines the semantics: R ranks x batch b  ==  one
process at batch R*b with gradient averaging, because InstanceNorm is per-sample and every loss is
a batch mean).

Gradients live in flat fp32 bucket buffers (each ``p.grad`` is a view with the parameter's own
packed strides), filled by autograd in backward order; a bucket's all-reduce (NCCL over
NVLink/NVSwitch, ReduceOp.AVG) is issued asynchronously from the post-accumulate-grad hook of its
last parameter, so communication overlaps the remaining backward kernels.  ``finish()`` waits
before the fused Adam sweep.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def _dense_view(t):
    """A contiguous view of a dense tensor whose dims are permuted in memory (conv weights keep the kernels'
    packed [taps][Cout][Cin] layout behind torch's logical shape); collectives need contiguous tensors."""
    if t.is_contiguous():
        return t
    perm = sorted(range(t.dim()), key=lambda i: -t.stride(i))
    v = t.permute(perm)
    if not v.is_contiguous():
        raise RuntimeError("parameter is not a dense permutation of a contiguous buffer")
    return v


def _slot(numel, align=64):
    """Elements a parameter occupies in its flat bucket: every view starts on a 256-byte boundary, so the wgrad
    kernels' 16-byte vector reductions (red.global.add.v4.f32, fused_wgrad mode) and the Adam sweep stay aligned
    whatever odd-sized tensors (1-element biases) precede it; the padding stays zero and rides along in the all-reduce."""
    return (numel + align - 1) // align * align


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work", "launched")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.work, self.launched = 0, None, False


class GradSync:
    """Bucketed, overlapped gradient averaging for a set of networks grouped into phases."""

    def __init__(self, phases, bucket_mb=32.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.use_avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.phases = {name: self._build(nets) for name, nets in phases.items()}
        self.active = None
        # ablation only (bench: how much of a step is exposed communication): gradients are NOT averaged
        self.skip = os.environ.get("MRA_DP_SKIP_ALLREDUCE", "0") == "1"
        self.stats = {"buckets": sum(len(b) for b in self.phases.values()), "allreduce_calls": 0, "late_launches": 0}

    def _build(self, nets):
        params = [p for net in nets for p in net.parameters()]
        buckets, cur, cur_n = [], [], 0
        for p in reversed(params):                      # grads become ready roughly in reverse order
            cur.append(p)
            cur_n += p.numel()
            if cur_n >= self.bucket_elems:
                buckets.append(cur)
                cur, cur_n = [], 0
        if cur:
            buckets.append(cur)
        out = []
        for plist in buckets:
            flat = torch.zeros(sum(_slot(p.numel()) for p in plist), dtype=torch.float32, device=plist[0].device)
            off = 0
            b = _Bucket(flat, plist)
            for p in plist:
                # a view with the parameter's (packed) strides so AccumulateGrad adds in place
                p.grad = torch.as_strided(flat, p.shape, p.stride(), off)
                off += _slot(p.numel())
                p.register_post_accumulate_grad_hook(self._make_hook(b))
            out.append(b)
        return out

    def _make_hook(self, bucket):
        def hook(param):
            if self.active is None or bucket.launched:
                return
            bucket.pending -= 1
            if bucket.pending == 0:
                self._launch(bucket)
        return hook

    def _launch(self, bucket):
        bucket.launched = True
        if self.world > 1 and not self.skip:
            op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM
            bucket.work = dist.all_reduce(bucket.flat, op=op, group=self.group, async_op=True)
            self.stats["allreduce_calls"] += 1

    def zero(self, phase):
        """zero_grad for a phase: one fill per flat bucket (every ``p.grad`` is a view into it)."""
        for b in self.phases[phase]:
            b.flat.zero_()

    def begin(self, phase):
        """Call after zero_grad(set_to_none=False) and before the phase's backward pass(es)."""
        self.active = phase
        for b in self.phases[phase]:
            b.pending = sum(1 for p in b.params if p.requires_grad)
            b.work, b.launched = None, False
            for p in b.params:                          # someone may have dropped the views (set_to_none)
                if p.grad is None or p.grad.data_ptr() < b.flat.data_ptr() or \
                        p.grad.data_ptr() >= b.flat.data_ptr() + b.flat.numel() * 4:
                    self._rebind(b)
                    break

    def _rebind(self, b):
        off = 0
        b.flat.zero_()
        for p in b.params:
            p.grad = torch.as_strided(b.flat, p.shape, p.stride(), off)
            off += _slot(p.numel())

    def finish(self, phase):
        """Issue any bucket whose parameters did not all receive a gradient, then wait."""
        for b in self.phases[phase]:
            lo, hi = b.flat.data_ptr(), b.flat.data_ptr() + b.flat.numel() * 4
            for p in b.params:                          # a replaced .grad would silently escape the averaging
                if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("GradSync: a gradient left its bucket during the backward pass")
            if not b.launched:
                self.stats["late_launches"] += 1        # no overlap for this one: not every hook of the bucket fired
                self._launch(b)
        for b in self.phases[phase]:
            if b.work is not None:
                b.work.wait()
                if not self.use_avg:
                    b.flat.mul_(1.0 / self.world)
                b.work = None
        self.active = None


def attach(model, bucket_mb=32.0, fused_wgrad=None):
    """Make a CycleGANModel data-parallel: broadcast rank 0's weights and buffers, then average
    gradients across ranks every step.  Returns the GradSync (also stored as ``model.grad_sync``)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if getattr(getattr(model, "opt", None), "norm", "instance") == "batch":
            raise NotImplementedError("data parallelism with norm='batch': batch statistics would be per rank "
                                      "(R ranks x b != batch R*b); use norm='instance' (the CycleGAN default)")
        for name in model.model_names:
            net = getattr(model, "net" + name)
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(_dense_view(t.data), src=0)
            # the broadcast wrote through .data: neither ._version nor data_ptr() moved, so the conv modules' cached
            # bf16 / transposed weight copies (and the Adam-maintained shadows) would still look valid -- a forward run
            # before attach() would leave ranks != 0 computing their first step with pre-broadcast weights
            for p in net.parameters():
                p._mra_epoch = getattr(p, "_mra_epoch", 0) + 1
                if hasattr(p, "_mra_shadow_tag"):
                    p._mra_shadow_tag = None
            for m in net.modules():
                if isinstance(getattr(m, "_cache", None), dict):
                    m._cache.clear()
    from . import networks3D
    # Default: every use of a weight hands autograd its own gradient, AccumulateGrad adds them into the bucket view.
    # fused_wgrad (MRA_DP_FUSED_WGRAD=1) keeps the single-process fusion instead: the wgrad kernels add every use
    # straight into the bucket view and autograd is handed nothing.  The engine still runs each parameter's
    # AccumulateGrad node -- and with it the post-accumulate hook the buckets listen to -- once all uses are done
    # (observed on torch 2.11; were it ever skipped, finish() launches the bucket, losing overlap, not correctness).
    # Default since it was measured at N = 2 (profiles/r02_dp2_variants_v2.txt: with graph replay 257.96 vs 259.77 ms on
    # resnet_9blocks, 58.35 vs 61.32 ms on unet_128, no late bucket launches); MRA_DP_FUSED_WGRAD=0 restores the
    # autograd accumulation.
    fused = os.environ.get("MRA_DP_FUSED_WGRAD", "1") == "1" if fused_wgrad is None else bool(fused_wgrad)
    for name in model.model_names:
        networks3D.set_fused_wgrad(getattr(model, "net" + name), fused)
    model.grad_sync = GradSync({"G": [model.netG_A, model.netG_B], "D": [model.netD_A, model.netD_B]},
                               bucket_mb=bucket_mb)
    return model.grad_sync
