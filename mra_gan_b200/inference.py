"""Sliding-window generator inference (reference test.py:96-185), windows sharded across ranks.

Same grid, clamping, scaling and averaging as the reference loop; windows are independent because
the generator keeps per-window instance statistics (train-mode norm: test.py never calls eval()),
so rank r takes windows r, r+R, ...; label/weight accumulators live on the device and one
reduce(sum) to rank 0 finishes the volume.
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
                             _local_only=False):
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
    for (i0, j0, k0) in grid[rank::world]:
        win = I.window_extract(vol, i0, j0, k0, patch, dtype)        # (1, px, py, pz, 1), scaled to [-1, 1]
        model.set_input(win.reshape(1, 1, *patch))
        model.test()
        pred = model.get_current_visuals()["fake_B"]
        I.window_accumulate(pred.reshape(1, *patch, 1).contiguous(), label, weight, i0, j0, k0)
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
