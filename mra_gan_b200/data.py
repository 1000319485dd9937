"""Device-side data path around the hot loop: what ``train.py:35-52,109-124`` and
``test.py:22-34,147-173`` do on the CPU with MONAI / SimpleITK, restated as a few tensor operations on the GPU so
that eight B200s are not starved by a ``workers=0`` loader.

* ``read_nifti`` -- minimal NIfTI-1 reader (single-file ``.nii`` / ``.nii.gz``; neither SimpleITK nor nibabel exists in
  this image).  Returns the voxel array in file order (x fastest), scaled by ``scl_slope / scl_inter`` when set.
* ``write_nifti`` -- the matching float32 writer for the translated volume (test.py:187-203).
* ``to_unit_range`` / ``from_unit_range`` -- the reference's intensity convention ``(x - 127.5) / 127.5`` (test.py:152)
  and its inverse (test.py:164).
* ``random_patch_pairs`` -- the role of ``RandCropByPosNegLabeld`` without labels: ``n`` random ``patch``-sized crops of
  an (A, B) volume pair taken on the device, returned as the ``(n, 1, D, H, W)`` fp32 batches ``set_input`` expects.

Plumbing only (torch indexing, no arithmetic kernels of its own); the window extraction / accumulation of the
sliding-window path lives in ``inference.py`` and ``csrc/misc.cuh``."""
import gzip
import struct

import numpy as np
import torch

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32}


def read_nifti(path):
    """-> (array of shape (x, y, z[, t]) as float32, voxel sizes (dx, dy, dz))."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError("%s: too short for a NIfTI-1 header" % path)
    endian = "<"
    if struct.unpack("<i", raw[0:4])[0] != 348:
        endian = ">"
        if struct.unpack(">i", raw[0:4])[0] != 348:
            raise ValueError("%s: not a NIfTI-1 file (sizeof_hdr != 348)" % path)
    if raw[344:348] not in (b"n+1\x00", b"ni1\x00"):
        raise ValueError("%s: bad NIfTI-1 magic %r" % (path, raw[344:348]))
    if raw[344:348] == b"ni1\x00":
        raise ValueError("%s: header/image pairs (.hdr/.img) are not supported" % path)
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(endian + "2h", raw[70:74])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset = int(struct.unpack(endian + "f", raw[108:112])[0])
    slope, inter = struct.unpack(endian + "2f", raw[112:120])
    if datatype not in _NIFTI_DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise ValueError("%s: bad dim[0] = %d" % (path, ndim))
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    dt = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(endian)
    count = int(np.prod(shape))
    if len(raw) < vox_offset + count * dt.itemsize:
        raise ValueError("%s: truncated voxel data" % path)
    arr = np.frombuffer(raw, dtype=dt, count=count, offset=vox_offset).reshape(shape, order="F").astype(np.float32)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0:
            arr = arr * np.float32(slope) + np.float32(inter)
    return np.ascontiguousarray(arr), tuple(float(p) for p in pixdim[1:4])


def write_nifti(path, array, voxel=(1.0, 1.0, 1.0)):
    """Minimal NIfTI-1 writer (float32, single file, ``.nii`` or ``.nii.gz``): what ``test.py:187-203`` needs to put
    the translated volume on disk.  ``array``: (x, y, z) numpy array or tensor."""
    arr = np.asarray(array.detach().cpu() if torch.is_tensor(array) else array, dtype=np.float32)
    if arr.ndim != 3:
        raise ValueError("write_nifti expects a 3-D volume")
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, arr.shape[0], arr.shape[1], arr.shape[2], 1, 1, 1, 1)
    struct.pack_into("<2h", hdr, 70, 16, 32)                       # datatype float32, bitpix
    struct.pack_into("<8f", hdr, 76, 1.0, float(voxel[0]), float(voxel[1]), float(voxel[2]), 0.0, 0.0, 0.0, 0.0)
    struct.pack_into("<f", hdr, 108, 352.0)                        # vox_offset
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)                    # scl_slope, scl_inter
    struct.pack_into("<2h", hdr, 252, 0, 1)                        # qform_code 0, sform_code 1 (scanner)
    struct.pack_into("<4f", hdr, 280, float(voxel[0]), 0.0, 0.0, 0.0)
    struct.pack_into("<4f", hdr, 296, 0.0, float(voxel[1]), 0.0, 0.0)
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, float(voxel[2]), 0.0)
    hdr[344:348] = b"n+1\x00"
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(arr.tobytes(order="F"))


def to_unit_range(x):
    """0..255 intensities -> the networks' [-1, 1] range (test.py:152)."""
    return (x - 127.5) / 127.5


def from_unit_range(y):
    """Inverse of to_unit_range (test.py:164)."""
    return y * 127.5 + 127.5


def random_patch_pairs(vol_a, vol_b, patch, n, generator=None, device=None, unit_range=True):
    """``n`` random crops of size ``patch`` (int or 3-tuple) from two volumes of equal shape.

    The crop origins come from ``generator`` (a CPU torch.Generator: reproducible, no device sync); the crops
    themselves are taken on ``device`` (default: where ``vol_a`` lives).  Volumes smaller than the patch along a
    dimension are edge-padded first, as the reference's loader pads (utils/NiftiDataset.py:391-503)."""
    if vol_a.shape != vol_b.shape or vol_a.dim() != 3:
        raise ValueError("random_patch_pairs expects two 3-D volumes of equal shape")
    p = (patch,) * 3 if isinstance(patch, int) else tuple(patch)
    dev = torch.device(device) if device is not None else vol_a.device
    a, b = vol_a.to(dev, torch.float32), vol_b.to(dev, torch.float32)
    pad = [max(0, p[i] - a.shape[i]) for i in range(3)]
    if any(pad):
        # F.pad pads the LAST dim first
        cfg = (0, pad[2], 0, pad[1], 0, pad[0])
        a = torch.nn.functional.pad(a[None, None], cfg, mode="replicate")[0, 0]
        b = torch.nn.functional.pad(b[None, None], cfg, mode="replicate")[0, 0]
    hi = [a.shape[i] - p[i] + 1 for i in range(3)]
    out_a = torch.empty((n, 1) + p, dtype=torch.float32, device=dev)
    out_b = torch.empty((n, 1) + p, dtype=torch.float32, device=dev)
    for i in range(n):
        o = [int(torch.randint(0, hi[d], (1,), generator=generator)) for d in range(3)]
        sl = tuple(slice(o[d], o[d] + p[d]) for d in range(3))
        out_a[i, 0].copy_(a[sl])
        out_b[i, 0].copy_(b[sl])
    if unit_range:
        out_a, out_b = to_unit_range(out_a), to_unit_range(out_b)
    return out_a, out_b
