"""ctypes loader for oracle/liboracle_kernels.so (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_kernels.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "csrc", "oracle_kernels.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_kernels.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64 = ctypes.c_int64
        L.orc_spgemm_csc_symbolic.restype = i64
        L.orc_spgemm_csc_symbolic.argtypes = [i64, i64, _i64p, _i64p, _i64p, _i64p, _i64p]
        L.orc_spgemm_csc_numeric.restype = ctypes.c_int
        L.orc_spgemm_csc_numeric.argtypes = [i64, i64, _i64p, _i64p, _f64p, _i64p, _i64p, _f64p,
                                             _i64p, _i64p, _f64p]
        L.orc_spmv_csc.restype = None
        L.orc_spmv_csc.argtypes = [i64, i64, _i64p, _i64p, _f64p, _f64p, _f64p]
        L.orc_mt_init.restype = None
        L.orc_mt_init.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.orc_mt_rand.restype = None
        L.orc_mt_rand.argtypes = [ctypes.c_void_p, i64, _f64p]
        L.orc_mt_sizeof.restype = ctypes.c_int
        L.orc_ax.restype = None
        L.orc_ax.argtypes = [_f64p, _f64p, _f64p, i64, i64, _f64p]
        L.orc_aty.restype = None
        L.orc_aty.argtypes = [_f64p, _f64p, _f64p, i64, i64, _f64p]
        _lib = L
    return _lib
