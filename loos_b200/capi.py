"""ctypes binding of include/loosb200.h -- the same stub a LOOS maintainer would
write for a Python front-end.  There is no fallback: if the
shared library is missing or no sm_100 device is present, calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libloosb200.so")

LAYOUT_XYZ, LAYOUT_PLANES = 0, 1
ENGINE_AUTO, ENGINE_TENSOR, ENGINE_FP64 = 0, 1, 2
ENGINES = {"auto": ENGINE_AUTO, "tensor": ENGINE_TENSOR, "fp64": ENGINE_FP64}

ERR_NO_DEVICE, ERR_CUDA, ERR_INVALID, ERR_SIZE_MISMATCH = -1, -2, -3, -4
ERR_RANGE, ERR_NOMEM, ERR_NUMERICAL, ERR_UNSUPPORTED = -5, -6, -7, -8


class Stats(C.Structure):
    _fields_ = [("max", C.c_double), ("sum", C.c_double), ("count", C.c_uint64)]


class LoosB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("loosb200 error %d: %s" % (code, text))
        self.code = code


_vp, _fp, _dp, _up = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
_sz, _i = C.c_size_t, C.c_int
_sp = C.POINTER(Stats)

# every symbol include/loosb200.h declares: name -> (restype, argtypes)
SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)   # loosb200_sink

SIGNATURES = {
    "loosb200_version": (C.c_char_p, []),
    "loosb200_strerror": (C.c_char_p, [_i]),
    "loosb200_last_error": (C.c_char_p, []),
    "loosb200_device_count": (_i, []),
    "loosb200_stage_frames": (_i, [C.POINTER(_vp), _vp, _sz, _sz, _up, _sz, _i, _i]),
    "loosb200_stage_frames_device": (_i, [C.POINTER(_vp), _vp, _sz, _sz, _up, _sz, _i, _i]),
    "loosb200_stage_begin": (_i, [C.POINTER(_vp), _sz, _sz, _up, _sz, _i, _i]),
    "loosb200_stage_append": (_i, [_vp, _vp, _sz]),
    "loosb200_stage_end": (_i, [_vp]),
    "loosb200_free": (None, [_vp]),
    "loosb200_trim": (None, []),
    "loosb200_frames_info": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    "loosb200_get_centroids": (_i, [_vp, _dp]),
    "loosb200_get_unit_log2": (_i, [_vp, C.POINTER(_i)]),
    "loosb200_rmsds_self": (_i, [_vp, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_self_device": (_i, [_vp, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_cross": (_i, [_vp, _vp, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_cross_device": (_i, [_vp, _vp, _vp, _sz, _sp, _i]),
    "loosb200_part_rows": (_i, [_sz, _i, _i, _up, C.POINTER(_sz)]),
    "loosb200_rmsds_self_part": (_i, [_vp, _i, _i, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_self_part_device": (_i, [_vp, _i, _i, _vp, _sz, _sp, _i]),
    "loosb200_rmsd2ref": (_i, [_vp, _vp, _dp, _dp, _dp, _dp]),
    "loosb200_rmsd2ref_iterative": (_i, [_vp, _vp, C.c_double, _i, _dp, _dp, _dp, C.POINTER(_i), _dp]),
    "loosb200_align_accumulate": (_i, [_vp, _vp, _dp, _dp, _dp]),
    "loosb200_rmsds_align": (_i, [_vp, _vp, _vp, _vp, _dp, _sz]),
    "loosb200_rmsds_align_engine": (_i, [_vp, _vp, _vp, _vp, _dp, _sz, _i]),
    "loosb200_share_cuts": (_i, [_sz, _i, _i, _dp, _up]),
    "loosb200_rmsds_self_text": (_i, [_vp, C.c_uint, SINK, _vp, _sp, _i]),
    "loosb200_rmsds_cross_text": (_i, [_vp, _vp, C.c_uint, SINK, _vp, _sp, _i]),
    "loosb200_last_timing": (_i, [_vp, _dp, _dp]),
    "loosb200_measure_int8_peak": (_i, [_i, C.c_double, _dp, _dp]),
    "loosb200_stage_frames_multi": (_i, [C.POINTER(_vp), C.POINTER(_i), _i, _vp, _sz, _sz, _up, _sz, _i]),
    "loosb200_replicate": (_i, [_vp, _i, C.POINTER(_vp)]),
    "loosb200_rmsds_self_multi": (_i, [C.POINTER(_vp), _i, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_self_multi_device": (_i, [C.POINTER(_vp), _i, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_cross_multi": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _vp, _sz, _sp, _i]),
    "loosb200_rmsds_self_multi_text": (_i, [C.POINTER(_vp), _i, C.c_uint, SINK, _vp, _sp, _i]),
    "loosb200_rmsds_cross_multi_text": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, C.c_uint, SINK, _vp, _sp, _i]),
    "loosb200_host_alloc": (_vp, [_sz]),
    "loosb200_host_free": (None, [_vp]),
}

_lib = None


def lib():
    """Load libloosb200.so (loudly: no library, no product)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m loos_b200.build` (needs nvcc); "
                              "loos_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code):
    if code < 0:
        L = lib()
        extra = L.loosb200_last_error().decode()
        raise LoosB200Error(code, L.loosb200_strerror(code).decode() + (": " + extra if extra and code == ERR_CUDA else ""))
    return code


def device_count():
    return lib().loosb200_device_count()
