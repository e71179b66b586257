"""ctypes binding of libstripenn_b200.so (C ABI: include/stripenn_b200.h).

The shared library is built in-tree by ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C stripenn_b200/csrc``).  There is NO fallback: if the library is missing, or no
B200 is present when a context is created, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# STRIPENN_B200_LIB: another build of the same library (kernel tuning experiments, tools/)
LIB_PATH = os.environ.get("STRIPENN_B200_LIB") or os.path.join(HERE, "lib", "libstripenn_b200.so")

NDIST = 400
NBRIGHT = 6
GRAY_IPP, GRAY_FMA2, GRAY_PLAIN = 0, 1, 2

OK, EINVAL, ENODEV, ECUDA, ENOMEM, ECAP, ESTATE, ERANGE = 0, -1, -2, -3, -4, -5, -6, -7


class Candidate(C.Structure):
    _fields_ = [("q", C.c_int32), ("frame", C.c_int32), ("bright", C.c_int32), ("ud", C.c_int32),
                ("k", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32),
                ("h", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("S", C.c_int32),
                ("pos1", C.c_int64), ("pos2", C.c_int64), ("pos3", C.c_int64), ("pos4", C.c_int64),
                ("total", C.c_double)]


CANDIDATE_DTYPE = np.dtype([("q", "<i4"), ("frame", "<i4"), ("bright", "<i4"), ("ud", "<i4"),
                            ("k", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"),
                            ("start", "<i4"), ("end", "<i4"), ("S", "<i4"), ("pos1", "<i8"),
                            ("pos2", "<i8"), ("pos3", "<i8"), ("pos4", "<i8"), ("total", "<f8")])
assert CANDIDATE_DTYPE.itemsize == C.sizeof(Candidate)

_P = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)

# every exported symbol of include/stripenn_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "snn_version": (C.c_int, []),
    "snn_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "snn_destroy": (None, [_P]),
    "snn_last_error": (C.c_char_p, [_P]),
    "snn_set_stream": (C.c_int, [_P, _P]),
    "snn_synchronize": (C.c_int, [_P]),
    "snn_load_chrom": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P,
                                 _P, _P, C.c_int]),
    "snn_load_chrom_async": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P,
                                       _P, _P, C.c_int]),
    "snn_load_chrom_device": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                        _P, _P, _P, _P, C.c_int]),
    "snn_unload_chrom": (C.c_int, [_P, C.c_int]),
    "snn_fetch": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "snn_heatmap": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_double, _P, _P]),
    "snn_quantiles": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "snn_quantiles_multi": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "snn_quantiles_select": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "snn_expected": (C.c_int, [_P, C.c_int, _P]),
    "snn_row_flags": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "snn_background": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P]),
    "snn_set_background": (C.c_int, [_P, C.c_int64, _P]),
    "snn_background_dev": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, C.c_int64]),
    "snn_set_background_dev": (C.c_int, [_P, C.c_int64, _P, _P]),
    "snn_device_alloc": (C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    "snn_device_free": (C.c_int, [_P, _P]),
    "snn_device_read": (C.c_int, [_P, _P, _P, C.c_int64]),
    "snn_extract": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_int, _P, C.c_int64, _i64p]),
    "snn_debug_stages": (C.c_int, [_P, C.c_int, C.c_double, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, _i32p, _P, _P, _P, _P, _P, _P, _P]),
    "snn_score": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "snn_score_multi": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "snn_set_mask": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64]),
    "snn_remove_redundant": (C.c_int, [C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "snn_stage_times": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), _f64p, _i64p]),
    "snn_reset_stage_times": (C.c_int, [_P]),
    "snn_enable_stage_timing": (C.c_int, [_P, C.c_int]),
    "snn_score_routing": (C.c_int, [_P, _i64p, C.c_int]),
}

_lib = None


class StripennB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("stripenn_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen the library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C stripenn_b200/csrc`.  stripenn_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """numpy array (or None) -> void*."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def check(rc, ctx=None):
    if rc == OK:
        return
    msg = load().snn_last_error(ctx)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == EINVAL and "Quantiles must be" in msg:
        raise ValueError(msg)
    raise StripennB200Error(rc, msg)
