"""ctypes binding of libssnmat.so (include/ssnamg_io.h): the C-side reader for the MAT-file level 5
inputs the reference scripts ``load`` (Class1/APD_SsN_Class1.m:27, Class2/APD_SsN_Class2.m:20).

``whos`` / ``read`` mirror MATLAB's ``whos('-file', ...)`` / ``load``; ``load_problem`` returns the
dictionary the drivers (``driver.py``) take.  No SciPy on this path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssnmat.so")

STATUS = {0: "SSN_MAT_OK", -1: "SSN_MAT_E_IO", -2: "SSN_MAT_E_FORMAT", -3: "SSN_MAT_E_ZLIB", -4: "SSN_MAT_E_NOMEM",
          -5: "SSN_MAT_E_INVALID", -6: "SSN_MAT_E_UNSUPPORTED"}


class MatError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{STATUS.get(code, code)}: {text}")
        self.code = code
        self.status = STATUS.get(code, str(code))


class Problem(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("c", C.POINTER(C.c_double)), ("r", C.POINTER(C.c_double)),
                ("l", C.POINTER(C.c_double)), ("p", C.POINTER(C.c_double)), ("q", C.POINTER(C.c_double)),
                ("gama", C.POINTER(C.c_double)), ("phi", C.POINTER(C.c_double)), ("mu", C.c_double)]


_vp, _pi64 = C.c_void_p, C.POINTER(C.c_int64)
SIGNATURES = {
    "ssn_mat_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "ssn_mat_close": (None, [_vp]),
    "ssn_mat_count": (C.c_int, [_vp]),
    "ssn_mat_name": (C.c_char_p, [_vp, C.c_int]),
    "ssn_mat_find": (C.c_int, [_vp, C.c_char_p]),
    "ssn_mat_dims": (C.c_int, [_vp, C.c_int, _pi64, _pi64]),
    "ssn_mat_read_double": (C.c_int, [_vp, C.c_int, _vp]),
    "ssn_problem_load": (C.c_int, [C.c_char_p, C.POINTER(Problem)]),
    "ssn_problem_free": (None, [C.POINTER(Problem)]),
    "ssn_mat_strerror": (C.c_char_p, [C.c_int]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build() (make -C csrc)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


def _check(st, what):
    if st != 0:
        raise MatError(st, f"{what}: {load().ssn_mat_strerror(st).decode()}")


class MatFile:
    """An opened MAT-file (context manager)."""

    def __init__(self, path):
        self._h = _vp()
        _check(load().ssn_mat_open(os.fsencode(path), C.byref(self._h)), path)
        self.path = path

    def close(self):
        if self._h:
            load().ssn_mat_close(self._h)
            self._h = _vp()

    __enter__ = lambda self: self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def whos(self):
        """[(name, (rows, cols), supported)] in file order."""
        lib, out = load(), []
        for i in range(lib.ssn_mat_count(self._h)):
            r, c = C.c_int64(), C.c_int64()
            st = lib.ssn_mat_dims(self._h, i, C.byref(r), C.byref(c))
            out.append((lib.ssn_mat_name(self._h, i).decode(), (r.value, c.value), st == 0))
        return out

    def __contains__(self, name):
        return load().ssn_mat_find(self._h, name.encode()) >= 0

    def read(self, name):
        """The variable as a (rows, cols) fp64 array (Fortran order, as MATLAB holds it)."""
        lib = load()
        i = lib.ssn_mat_find(self._h, name.encode())
        if i < 0:
            raise KeyError(name)
        r, c = C.c_int64(), C.c_int64()
        _check(lib.ssn_mat_dims(self._h, i, C.byref(r), C.byref(c)), name)
        out = np.empty((r.value, c.value), dtype=np.float64, order="F")
        _check(lib.ssn_mat_read_double(self._h, i, out.ctypes.data_as(_vp)), name)
        return out


def load_problem(path):
    """``ssn_problem_load``: the variables of one bundled problem file as the driver dictionary
    (``gama`` collapses to ``np.inf`` when every entry is +Inf, which is how the drivers select the
    unbounded kernels)."""
    pb = Problem()
    lib = load()
    _check(lib.ssn_problem_load(os.fsencode(path), C.byref(pb)), path)
    try:
        m, n = int(pb.m), int(pb.n)
        cp = lambda ptr, k: np.ctypeslib.as_array(ptr, shape=(k,)).copy()
        out = {"m": m, "n": n, "c": cp(pb.c, m * n), "r": cp(pb.r, n), "l": cp(pb.l, m), "p": cp(pb.p, m), "q": cp(pb.q, n)}
        if pb.gama:
            g = cp(pb.gama, m * n)
            out["gama"] = np.inf if np.all(np.isposinf(g)) else g
        else:
            out["gama"] = np.inf
        if pb.phi:
            out["phi"] = cp(pb.phi, m * n)
        if not np.isnan(pb.mu):
            out["mu"] = float(pb.mu)
        return out
    finally:
        lib.ssn_problem_free(C.byref(pb))
