"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every symbol that
include/mra_gan_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from mra_gan_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mra_gan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mra_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), "missing export: " + s
    assert sorted(_lib.exported_symbols()) == syms, "ctypes binding table and header disagree"
    L = _lib.lib()
    assert L.mra_version() >= 100
    assert L.mra_debug_launch_count() == 0


def test_host_side_argument_validation():
    """Error plumbing across the ABI: negative code + message, nothing throws, no GPU touched."""
    L = _lib.lib()
    d = _lib.ConvDesc()
    d.n, d.cin, d.cout, d.k, d.stride = 1, 4, 4, 3, 3
    buf = (ctypes.c_int32 * 16)()
    assert L.mra_conv_plan_describe(ctypes.byref(d), 0, buf, 16) < 0
    assert b"stride" in L.mra_last_error()
    d.stride, d.pad = 1, 1
    d.din = d.hin = d.win = 8
    d.dout = d.hout = d.wout = 7                # inconsistent output size
    assert L.mra_conv_plan_describe(ctypes.byref(d), 0, buf, 16) < 0
    assert b"expected 8" in L.mra_last_error()
    d.dout = d.hout = d.wout = 8
    assert L.mra_conv_plan_describe(ctypes.byref(d), 0, buf, 16) == -2     # buffer too small


def test_null_operands_and_bad_dtype_are_rejected_on_the_host():
    """The conv entry points return an error code for a NULL mandatory operand or an unknown dtype before any CUDA call
    (this box has no GPU: a rejected call must not have needed one); nothing is launched."""
    L = _lib.lib()
    d = _lib.ConvDesc()
    d.n, d.cin, d.cout, d.k, d.stride, d.pad = 1, 4, 4, 3, 1, 1
    d.din = d.hin = d.win = d.dout = d.hout = d.wout = 8
    n0 = L.mra_debug_launch_count()
    one = ctypes.c_void_p(0x1000)                # never dereferenced: the calls fail on the NULL next to it
    assert L.mra_conv3d_fprop(ctypes.byref(d), None, one, None, one, None, None, 0, None) < 0
    assert b"mra_conv3d_fprop: null operand" in L.mra_last_error()
    assert L.mra_conv3d_fprop(ctypes.byref(d), one, one, None, None, None, None, 0, None) < 0
    assert L.mra_conv3d_dgrad(ctypes.byref(d), one, None, one, None, 0, None) < 0
    assert b"mra_conv3d_dgrad: null operand" in L.mra_last_error()
    assert L.mra_conv3d_dgrad_nstats(ctypes.byref(d), None, one, one, one, 0, 0.0, one, None, 0, None) < 0
    assert b"mra_conv3d_dgrad_nstats: null operand" in L.mra_last_error()
    assert L.mra_conv3d_wgrad(ctypes.byref(d), one, None, one, None, None, 0, None) < 0
    assert L.mra_conv3d_wgrad(ctypes.byref(d), None, one, one, None, None, 0, None) < 0      # dw asked for without x
    assert b"mra_conv3d_wgrad: null operand" in L.mra_last_error()
    d.dtype = 7
    assert L.mra_conv3d_fprop(ctypes.byref(d), one, one, None, one, None, None, 0, None) < 0
    assert b"unknown dtype 7" in L.mra_last_error()
    assert L.mra_conv3d_fprop(None, one, one, None, one, None, None, 0, None) < 0
    assert b"null conv descriptor" in L.mra_last_error()
    assert L.mra_debug_launch_count() == n0


def test_sass_contains_blackwell_tensor_core_and_tma_instructions():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass
