"""Host-side data path: minimal NIfTI reader, patch sampling, optimiser-state checkpoints."""
import gzip
import os
import random
import struct

import numpy as np
import pytest
import torch

from mra_gan_b200 import data
from mra_gan_b200 import networks3D as N3
from mra_gan_b200 import ops
from mra_gan_b200.models import create_model
from oracle import functional as OF
from oracle import ops_ref as R
from oracle.ref_import import make_opt


def _write_nifti(path, arr, dtype_code, slope=1.0, inter=0.0, endian="<", gz=False):
    hdr = bytearray(352)
    struct.pack_into(endian + "i", hdr, 0, 348)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into(endian + "8h", hdr, 40, *dim)
    struct.pack_into(endian + "2h", hdr, 70, dtype_code, arr.dtype.itemsize * 8)
    struct.pack_into(endian + "8f", hdr, 76, 1.0, 0.5, 0.5, 1.25, 0, 0, 0, 0)
    struct.pack_into(endian + "f", hdr, 108, 352.0)
    struct.pack_into(endian + "2f", hdr, 112, slope, inter)
    hdr[344:348] = b"n+1\x00"
    body = arr.astype(arr.dtype.newbyteorder(endian)).tobytes(order="F")
    with (gzip.open if gz else open)(path, "wb") as f:
        f.write(bytes(hdr) + body)


@pytest.mark.parametrize("code,dt,endian,gz", [(4, np.int16, "<", False), (16, np.float32, ">", True), (2, np.uint8, "<", True)])
def test_read_nifti_round_trip(tmp_path, code, dt, endian, gz):
    rng = np.random.default_rng(3)
    arr = (rng.random((7, 5, 4)) * 200).astype(dt)
    path = str(tmp_path / ("v.nii.gz" if gz else "v.nii"))
    _write_nifti(path, arr, code, slope=2.0, inter=-3.0, endian=endian, gz=gz)
    got, vox = data.read_nifti(path)
    assert got.shape == (7, 5, 4) and got.dtype == np.float32
    np.testing.assert_allclose(got, arr.astype(np.float32) * 2.0 - 3.0, rtol=0, atol=1e-4)
    assert vox == (0.5, 0.5, 1.25)


def test_write_nifti_round_trip(tmp_path):
    vol = torch.rand(9, 6, 5) * 255
    for name in ("o.nii", "o.nii.gz"):
        p = str(tmp_path / name)
        data.write_nifti(p, vol, voxel=(0.4, 0.4, 0.8))
        got, vox = data.read_nifti(p)
        assert np.array_equal(got, vol.numpy()) and np.allclose(vox, (0.4, 0.4, 0.8))


def test_read_nifti_rejects_garbage(tmp_path):
    p = str(tmp_path / "bad.nii")
    open(p, "wb").write(b"\x00" * 400)
    with pytest.raises(ValueError):
        data.read_nifti(p)


def test_random_patch_pairs_reproducible_and_aligned():
    g = torch.Generator().manual_seed(5)
    a = torch.arange(20 * 18 * 16, dtype=torch.float32).reshape(20, 18, 16)
    b = a * 2 + 1
    pa, pb = data.random_patch_pairs(a, b, (8, 6, 4), 3, generator=g, unit_range=False)
    assert pa.shape == (3, 1, 8, 6, 4) and pb.shape == pa.shape
    assert torch.equal(pb, pa * 2 + 1)                      # the same window of both volumes
    g2 = torch.Generator().manual_seed(5)
    qa, _ = data.random_patch_pairs(a, b, (8, 6, 4), 3, generator=g2, unit_range=False)
    assert torch.equal(pa, qa)
    # every crop is a contiguous window of the volume
    for i in range(3):
        o = int(pa[i, 0, 0, 0, 0])
        d, rem = divmod(o, 18 * 16)
        h, w = divmod(rem, 16)
        assert torch.equal(pa[i, 0], a[d:d + 8, h:h + 6, w:w + 4])
    # small volumes are edge-padded up to the patch size; unit range conversion is the reference's
    sa, sb = data.random_patch_pairs(a[:4], b[:4], 8, 1, generator=g)
    assert sa.shape == (1, 1, 8, 8, 8)
    assert torch.allclose(data.from_unit_range(data.to_unit_range(a)), a)


def test_optimizer_state_checkpoint_round_trip(tmp_path):
    """save_networks() also writes the Adam moments; a model restored with continue_train takes the same next step."""
    ops.set_impl(R.RefImpl(torch.float32))
    try:
        N3.set_default_compute_dtype(torch.float32)
        opt = make_opt(ngf=4, ndf=4, pool_size=0, checkpoints_dir=str(tmp_path), gpu_ids=-1)
        random.seed(1)
        torch.manual_seed(1)
        m = create_model(opt)
        m.setup(opt)
        A, B = OF.synthetic_patches(1, 32, seed=1)
        m.set_input([A, B])
        m.optimize_parameters()
        m.save_networks("latest")
        assert os.path.exists(os.path.join(m.save_dir, "latest_optim.pth"))
        A2, B2 = OF.synthetic_patches(1, 32, seed=2)
        m.set_input([A2, B2])
        m.optimize_parameters()
        want = {k: v.detach().clone() for k, v in m.netG_A.state_dict().items()}

        opt2 = make_opt(ngf=4, ndf=4, pool_size=0, checkpoints_dir=str(tmp_path), gpu_ids=-1, continue_train=True)
        m2 = create_model(opt2)
        m2.setup(opt2)
        p0 = next(m2.netG_A.parameters())
        st = m2.optimizer_G.state[p0]
        assert st["step"] == 1 and st["exp_avg"].stride() == p0.stride()
        m2.set_input([A2, B2])
        m2.optimize_parameters()
        for k, v in m2.netG_A.state_dict().items():
            assert torch.allclose(v, want[k], rtol=1e-5, atol=1e-7), k
    finally:
        ops.set_impl(None)
        N3.set_default_compute_dtype(torch.bfloat16)
