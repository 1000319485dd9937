"""ctypes binding of libmra_b200.so (the C ABI in include/mra_gan_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc (cross-compiles without
a GPU), and if that fails the import error is raised to the caller.
"""
import ctypes as C
import os
import threading

from . import build as _build

MRA_F32, MRA_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
LOSS_L1, LOSS_MSE_CONST, LOSS_BCE_CONST = 0, 1, 2
CONV_FORCE_NAIVE, CONV_ACCUMULATE, CONV_WS_REUSE = 1, 2, 4


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n", "cin", "cout", "din", "hin", "win", "dout", "hout", "wout", "k", "stride", "pad",
                 "transposed", "dtype", "act")] + [("slope", C.c_float), ("flags", C.c_int32)]


class NormDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("c", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("pad", C.c_int32), ("act", C.c_int32), ("slope", C.c_float), ("res_pad", C.c_int32),
                ("dtype", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
                ("use_running", C.c_int32)]


class AdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("shadow", C.c_void_p), ("numel", C.c_int64)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGNATURES = {
    "mra_version": ([], C.c_int),
    "mra_last_error": ([], C.c_char_p),
    "mra_debug_tc_error": ([_I], C.c_int),
    "mra_debug_launch_count": ([], C.c_longlong),
    "mra_debug_counters": ([C.POINTER(C.c_ulonglong), _I], C.c_int),
    "mra_conv3d_fprop": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, C.c_size_t, _P], C.c_int),
    "mra_conv3d_dgrad": ([C.POINTER(ConvDesc), _P, _P, _P, _P, C.c_size_t, _P], C.c_int),
    "mra_conv3d_dgrad_nstats_supported": ([C.POINTER(ConvDesc)], C.c_int),
    "mra_conv3d_dgrad_nstats": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _I, _F, _P, _P, C.c_size_t, _P], C.c_int),
    "mra_conv3d_wgrad": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, C.c_size_t, _P], C.c_int),
    "mra_conv3d_workspace_size": ([C.POINTER(ConvDesc), _I], C.c_size_t),
    "mra_conv3d_uses_tensor_cores": ([C.POINTER(ConvDesc), _I], C.c_int),
    "mra_conv3d_lowering": ([C.POINTER(ConvDesc)], C.c_int),
    "mra_pack_weight_t": ([_P, _I, _P, _I, _I, _I, _I, _P], C.c_int),
    "mra_convert": ([_P, _I, _P, _I, _L, _P], C.c_int),
    "mra_inorm_stats": ([C.POINTER(NormDesc), _P, _P, _P], C.c_int),
    "mra_inorm_act_pad_fwd": ([C.POINTER(NormDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "mra_inorm_act_pad_bwd": ([C.POINTER(NormDesc), _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "mra_inorm_act_pad_bwd_stats": ([C.POINTER(NormDesc), _P, _P, _P, _P, _P, _P], C.c_int),
    "mra_inorm_act_pad_bwd_apply": ([C.POINTER(NormDesc), _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "mra_act_fwd": ([_P, _P, _L, _I, _F, _I, _P], C.c_int),
    "mra_act_bwd": ([_P, _P, _P, _L, _I, _F, _I, _P], C.c_int),
    "mra_cat2_act_fwd": ([_P, _P, _P, _L, _I, _I, _I, _F, _I, _P], C.c_int),
    "mra_cat2_act_bwd": ([_P, _P, _P, _P, _L, _I, _I, _I, _F, _I, _P], C.c_int),
    "mra_mask_scale": ([_P, _P, _P, _L, _F, _I, _P], C.c_int),
    "mra_reppad_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "mra_reppad_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "mra_loss_fwd": ([_I, _P, _P, _F, _L, _I, _P, _P], C.c_int),
    "mra_loss_bwd": ([_I, _P, _P, _F, _L, _I, _P, _F, _P, _P], C.c_int),
    "mra_corr_sums": ([_P, _P, _L, _I, _P, _P], C.c_int),
    "mra_adam_multi": ([C.POINTER(AdamTensor), _I, _F, _F, _F, _F, _I, _P], C.c_int),
    "mra_adam_multi_dev": ([C.POINTER(AdamTensor), _I, _P, _P], C.c_int),
    "mra_adam_advance": ([_P, _P, _P], C.c_int),
    "mra_window_extract": ([_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P], C.c_int),
    "mra_window_accumulate": ([_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "mra_window_finalize": ([_P, _P, _L, _P], C.c_int),
    "mra_conv_plan_describe": ([C.POINTER(ConvDesc), _I, C.POINTER(C.c_int32), _I], C.c_int),
    "mra_debug_schedule": ([C.POINTER(ConvDesc), _I, _I, _I, C.POINTER(C.c_int32), _I], C.c_int),
}

_lock = threading.Lock()
_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded library (built on first use).  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = _build.LIB
            if _build.needs_build():
                path = _build.build()
            handle = C.CDLL(path)
            for name, (argtypes, restype) in _SIGNATURES.items():
                fn = getattr(handle, name)          # AttributeError if the symbol is missing
                fn.argtypes, fn.restype = argtypes, restype
            _lib = handle
    return _lib


class MraError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().mra_last_error()
        raise MraError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
