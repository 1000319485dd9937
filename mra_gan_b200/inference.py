"""Sliding-window generator inference (reference test.py:96-185), windows sharded across ranks.

Same grid, clamping, scaling and averaging as the reference loop; windows are independent because
the generator keeps per-window instance statistics (train-mode norm: test.py never calls eval()),
so rank r takes windows r, r+R, ...; label/weight accumulators live on the device and one
reduce(sum) to rank 0 finishes the volume.

For the same reason a rank may push several of its windows through the generator as ONE batch
(``windows_per_pass``, default 4): every sample of a batch gets its own statistics, so the result of a window
does not depend on its batch mates, while the persistent conv kernels fill their last round of tiles
(256 -> 256 k3 at 32^3: 128 CTA-pair tiles on 74 pairs at batch 1 = 2 rounds for 1.73 rounds of work, 256 tiles at
batch 2 = 3.5 for 3.46; config 5 on one B200: 162.6 / 150.0 / 143.4 ms per volume at 1 / 2 / 4 windows per pass,
profiles/r02_infer_windows_per_pass.txt).  The accumulation order of overlapping windows is unchanged (windows are added in grid
order), so the sums are those of the one-window-at-a-time loop.
"""
import math

import torch
import torch.distributed as dist

from . import ops


def window_grid(shape, patch, stride_inplane, stride_layer):
    """Window corners in the reference's i, j, k order with the last window clamped to the edge
    (test.py:111-145)."""
    X, Y, Z = shape
    px, py, pz = patch
    counts = [int(math.ceil((X - px) / float(stride_inplane))) + 1,
              int(math.ceil((Y - py) / float(stride_inplane))) + 1,
              int(math.ceil((Z - pz) / float(stride_layer))) + 1]
    grid = []
    for i in range(counts[0]):
        i0 = min(i * stride_inplane, X - px)
        for j in range(counts[1]):
            j0 = min(j * stride_inplane, Y - py)
            for k in range(counts[2]):
                k0 = min(k * stride_layer, Z - pz)
                grid.append((i0, j0, k0))
    return grid


@torch.no_grad()
def sliding_window_inference(model, volume, patch, stride_inplane, stride_layer, rank=0, world=1, dtype=None,
                             _local_only=False, windows_per_pass=4):
    """``model``: a TestModel (set_input / test / get_current_visuals, like test.py:158-161).
    ``volume``: float32 (X, Y, Z) tensor on the 0..255 scale, each dim >= patch.
    Returns the (X, Y, Z) float32 result on rank 0 (other ranks get their partial sum's buffer)."""
    I = ops.impl()
    dev = model.device
    vol = torch.as_tensor(volume, dtype=torch.float32).to(dev).contiguous()
    padded = vol.shape[2] % 2 != 0
    if padded:                                                       # test.py:98-103
        vol = torch.cat([vol, vol[:, :, -1:]], 2).contiguous()
    label = torch.zeros_like(vol)
    weight = torch.zeros_like(vol)
    dtype = dtype or torch.float32
    grid = window_grid(tuple(vol.shape), patch, stride_inplane, stride_layer)
    mine = grid[rank::world]
    step = max(1, int(windows_per_pass))
    net = getattr(model, "netG", None)
    if net is not None and any(type(mod).__name__ == "BatchNorm3d" for mod in net.modules()):
        step = 1                    # train-mode batch statistics couple the samples of a batch: keep the reference's batch of 1
    for b0 in range(0, len(mine), step):
        group = mine[b0:b0 + step]
        wins = [I.window_extract(vol, i0, j0, k0, patch, dtype) for (i0, j0, k0) in group]   # (1, px, py, pz, 1) in [-1, 1]
        win = wins[0] if len(wins) == 1 else torch.cat(wins, 0)
        model.set_input(win.reshape(len(group), 1, *patch))
        model.test()
        pred = model.get_current_visuals()["fake_B"]                 # (len(group), 1, px, py, pz)
        for b, (i0, j0, k0) in enumerate(group):
            I.window_accumulate(pred[b].reshape(1, *patch, 1).contiguous(), label, weight, i0, j0, k0)
    if _local_only:                 # test hook: this rank's partial sums, before the cross-rank reduce
        return label, weight
    if world > 1:
        dist.reduce(label, dst=0, op=dist.ReduceOp.SUM)
        dist.reduce(weight, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        I.window_finalize(label, weight)                             # label / weight + 0.01  (test.py:178)
    if padded:
        label = label[:, :, :-1]
    return label
