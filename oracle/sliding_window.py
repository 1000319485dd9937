"""numpy restatement of the reference's sliding-window inference loop (TEST INFRASTRUCTURE).

Follows /root/reference/test.py:96-185 line by line (SimpleITK I/O around it is out of scope):
odd-z edge pad (:98-103), window grid (:111-113), last-window clamping (:125-138), input scaling
(:152), output un-scaling (:164), accumulate + count (:172-173), normalise + 0.01 (:178), drop
the pad slice (:181-182).
"""
import math

import numpy as np


def window_grid(shape, patch, stride_inplane, stride_layer):
    """Window start/end indices in the reference's i, j, k order (test.py:111-145)."""
    X, Y, Z = shape
    px, py, pz = patch
    inum = int(math.ceil((X - px) / float(stride_inplane))) + 1
    jnum = int(math.ceil((Y - py) / float(stride_inplane))) + 1
    knum = int(math.ceil((Z - pz) / float(stride_layer))) + 1
    out = []
    for i in range(inum):
        for j in range(jnum):
            for k in range(knum):
                i0 = i * stride_inplane
                if i0 + px > X:
                    i0 = X - px
                j0 = j * stride_inplane
                if j0 + py > Y:
                    j0 = Y - py
                k0 = k * stride_layer
                if k0 + pz > Z:
                    k0 = Z - pz
                out.append((i0, i0 + px, j0, j0 + py, k0, k0 + pz))
    return out


def sliding_window_inference(image_np, generator_fn, patch, stride_inplane, stride_layer):
    """``image_np``: float32 (X, Y, Z) on the 0..255 scale, already >= patch in every dim.
    ``generator_fn``: (1,1,px,py,pz) float32 array -> (px,py,pz) float32 array (train-mode G)."""
    image_np = np.asarray(image_np)
    label_np = np.zeros(image_np.shape, np.float32)
    padded = image_np.shape[2] % 2 != 0
    if padded:
        image_np = np.pad(image_np, ((0, 0), (0, 0), (0, 1)), "edge")
        label_np = np.pad(label_np, ((0, 0), (0, 0), (0, 1)), "edge")
    weight_np = np.zeros(label_np.shape)
    for (i0, i1, j0, j1, k0, k1) in window_grid(image_np.shape, patch, stride_inplane, stride_layer):
        batch = (image_np[i0:i1, j0:j1, k0:k1][np.newaxis] - 127.5) / 127.5
        pred = generator_fn(batch[np.newaxis].astype(np.float32))
        pred = pred * 127.5 + 127.5
        label_np[i0:i1, j0:j1, k0:k1] += pred
        weight_np[i0:i1, j0:j1, k0:k1] += 1.0
    label_np = np.float32(label_np) / np.float32(weight_np) + 0.01
    if padded:
        label_np = label_np[:, :, :-1]
    return label_np
