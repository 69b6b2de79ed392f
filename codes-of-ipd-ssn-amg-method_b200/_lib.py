"""ctypes binding of libssnamg.so (the C ABI of include/ssnamg.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present, every
operator raises.  The library lives in-tree next to this file (built by ``__graft_entry__.build``
or ``make -C csrc``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSN_LIB_PATH") or os.path.join(_HERE, "libssnamg.so")     # SSN_LIB_PATH: a development build

STATUS = {
    0: "SSN_OK", -1: "SSN_E_CUDA", -2: "SSN_E_INVALID", -3: "SSN_E_PQ_ZERO", -4: "SSN_E_BIGPH_FNODE",
    -5: "SSN_E_PCG_NF", -6: "SSN_E_NOT_SQUARE", -7: "SSN_E_CF_PARTITION", -8: "SSN_E_COARSEN_STALL",
    -9: "SSN_E_NOT_BIGRAPH", -10: "SSN_E_UNSUPPORTED", -11: "SSN_E_TOO_LARGE", -12: "SSN_E_NOT_SPD",
    -13: "SSN_E_NO_HIERARCHY", -14: "SSN_E_ASATZ_DIM",
}


class SsnError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{STATUS.get(code, code)}: {text}")
        self.code = code
        self.status = STATUS.get(code, str(code))


class CSR(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64),
                ("rowptr_dev", C.c_void_p), ("colidx_dev", C.c_void_p), ("val_dev", C.c_void_p)]


class AmgOptions(C.Structure):
    _fields_ = [("retol", C.c_double), ("bigph", C.c_int32), ("maxit", C.c_int32), ("theta", C.c_double),
                ("smoth", C.c_int32), ("cycle", C.c_int32), ("isnsp", C.c_int32), ("inter", C.c_int32),
                ("fnode", C.c_int32), ("guess_dev", C.c_void_p)]


class PcgOptions(C.Structure):
    _fields_ = [("retol", C.c_double), ("maxit", C.c_int32), ("precd", C.c_int32), ("nf", C.c_int32),
                ("guess_dev", C.c_void_p)]


class ProbData(C.Structure):
    _fields_ = [("bk1", C.c_double), ("tk", C.c_double), ("m", C.c_int64), ("n", C.c_int64),
                ("p_dev", C.c_void_p), ("q_dev", C.c_void_p), ("t_dev", C.c_void_p),
                ("H0", C.POINTER(CSR)), ("z_dev", C.c_void_p), ("s_dev", C.c_void_p), ("phi_dev", C.c_void_p)]


class ApdOptions(C.Structure):
    _fields_ = [("inner_solver", C.c_int32), ("maxit", C.c_int32), ("KKT_Tol", C.c_double), ("warm_maxit", C.c_int32),
                ("max_outer", C.c_int32), ("max_seconds", C.c_double), ("verbose", C.c_int32),
                ("amg", C.POINTER(AmgOptions)), ("pcg", C.POINTER(PcgOptions))]


class ApdResult(C.Structure):
    _fields_ = [("outer_its", C.c_int32), ("converged", C.c_int32), ("rel_kkt", C.c_double), ("objective", C.c_double),
                ("ssn_steps", C.c_int32), ("ls_trials", C.c_int32), ("ls_passes", C.c_int32), ("amg_calls", C.c_int32),
                ("warmup_s", C.c_double), ("loop_s", C.c_double), ("solve_s", C.c_double), ("asat_s", C.c_double),
                ("plan_s", C.c_double), ("hist_len", C.c_int32), ("steps_len", C.c_int64)]


# every symbol include/ssnamg.h declares, with its ctypes signature
_vp, _i64, _i32, _dbl, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int
_pcsr, _pint, _pdbl, _pi64 = C.POINTER(CSR), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)
SIGNATURES = {
    "ssn_create": (_int, [C.POINTER(_vp), _int]),
    "ssn_destroy": (_int, [_vp]),
    "ssn_default_ctx_acquire": (_int, [C.POINTER(_vp)]),
    "ssn_default_ctx_release": (_int, []),
    "ssn_default_ctx_refcount": (_int, []),
    "ssn_last_error": (C.c_char_p, [_vp]),
    "ssn_set_stream": (_int, [_vp, _vp]),
    "ssn_synchronize": (_int, [_vp]),
    "ssn_version": (_int, []),
    "ssn_launch_count": (_i64, [_vp]),
    "ssn_profile_enable": (_int, [_vp, _int]),
    "ssn_set_dense_tail": (_int, [_vp, _int, _int]),
    "ssn_set_persistent": (_int, [_vp, _int]),
    "ssn_set_device_setup": (_int, [_vp, _int]),
    "ssn_set_fused_setup": (_int, [_vp, _int]),
    "ssn_set_cluster_solve": (_int, [_vp, _int]),
    "ssn_set_spgemm_slab_limit": (_int, [_vp, _i64]),
    "ssn_debug_barrier_bench": (_int, [_vp, _int, _int, _pdbl]),
    "ssn_kernel_timer": (_int, [_vp, _int]),
    "ssn_kernel_timer_read": (_int, [_vp, _pdbl, _pi64]),
    "ssn_profile_dump": (C.c_char_p, [_vp]),
    "ssn_debug_cycles": (_int, [_vp, _vp, _int]),
    "ssn_debug_cycles_persist": (_int, [_vp, _vp, _int]),
    "ssn_rng_reset": (_int, [_vp, C.c_uint32]),
    "ssn_rng_drawn": (_i64, [_vp]),
    "ssn_rand": (_int, [_vp, _i64, _vp]),
    "ssn_malloc": (_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "ssn_free": (_int, [_vp, _vp]),
    "ssn_memcpy_h2d": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "ssn_memcpy_d2h": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "ssn_memcpy_d2d": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "ssn_csr_free": (_int, [_vp, _pcsr]),
    "ssn_csr_upload": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _pcsr]),
    "ssn_csr_download": (_int, [_vp, _pcsr, _vp, _vp, _vp]),
    "ssn_ax": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ssn_ax_host": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ssn_aty": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ssn_aty_host": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ssn_prox_residual": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _vp, _vp, _vp,
                                 _pdbl, _pi64]),
    "ssn_prox_residual_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "ssn_prox_residual_pot": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _vp, _pdbl, _pi64]),
    "ssn_ssn_step_class1": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _dbl, _int, _vp, _vp, _vp, _vp]),
    "ssn_ssn_step_class1_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _dbl, _int, _vp, _vp, _vp, _vp]),
    "ssn_prox_trials": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp]),
    "ssn_prox_trials_lin": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _int, _int, _vp]),
    "ssn_warmup_class1": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _dbl, _int, _vp, _vp]),
    "ssn_warm_stage": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _dbl, _dbl, _dbl, _dbl,
                              _vp, _vp]),
    "ssn_apd_begin": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _vp]),
    "ssn_apd_end": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _dbl, _vp, _vp, _vp, _pdbl, _pdbl]),
    "ssn_trial_vectors": (_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _int, _int, _vp, _vp]),
    "ssn_linesearch": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _dbl, _dbl, _dbl, _int, _dbl, _dbl,
                              _int, _vp, C.POINTER(_int), _pdbl, _pdbl, C.POINTER(_int)]),
    "ssn_apd_ssn_class1": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _dbl, C.POINTER(ApdOptions), _vp, _vp,
                                  C.POINTER(ApdResult), _vp, _vp, _vp, _vp, _vp, _i64]),
    "ssn_apd_ssn_class1_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _dbl, C.POINTER(ApdOptions), _vp, _vp,
                                       C.POINTER(ApdResult), _vp, _vp, _vp, _vp, _vp, _i64]),
    "ssn_warmup_class2": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _int, _vp, _vp]),
    "ssn_apd_begin_pot": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "ssn_apd_end_pot": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    "ssn_apd_ssn_class2": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, C.POINTER(ApdOptions), _vp, _vp,
                                  C.POINTER(ApdResult), _vp, _vp, _vp, _vp, _i64]),
    "ssn_apd_ssn_class2_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, C.POINTER(ApdOptions), _vp, _vp,
                                       C.POINTER(ApdResult), _vp, _vp, _vp, _vp, _i64]),
    "ssn_ssn_step_class2": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "ssn_ssn_step_class2_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "ssn_amg4pot_str": (_int, [_vp, C.POINTER(ProbData), C.POINTER(AmgOptions), _int, _vp, _pint, _pdbl, _pint]),
    "ssn_asat": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _pcsr]),
    "ssn_asat_host": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _pcsr]),
    "ssn_active_coo": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, C.POINTER(_vp), _pi64]),
    "ssn_asat_coo": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _pcsr]),
    "ssn_asatz": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ssn_invaat": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    "ssn_invhht": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _vp]),
    "ssn_strength": (_int, [_vp, _pcsr, _int, _pcsr]),
    "ssn_mis_set": (_int, [_vp, _pcsr, _dbl, _vp, _vp, _pcsr]),
    "ssn_cf_split": (_int, [_vp, _pcsr, _vp, _vp]),
    "ssn_transfer": (_int, [_vp, _pcsr, C.POINTER(AmgOptions), _int, _pcsr, _pcsr, _pcsr, _vp]),
    "ssn_amg_setup": (_int, [_vp, _pcsr, C.POINTER(AmgOptions), _pint]),
    "ssn_amg_level": (_int, [_vp, _int, _pcsr, _pcsr]),
    "ssn_amg_clear": (_int, [_vp]),
    "ssn_mg_vcycle": (_int, [_vp, _vp, _int, _int, _vp]),
    "ssn_mg_wcycle": (_int, [_vp, _vp, _int, _int, _vp]),
    "ssn_class_amg": (_int, [_vp, _pcsr, _vp, C.POINTER(AmgOptions), _int, _vp, _pint, _pdbl, _vp, _vp, _pint]),
    "ssn_pcg": (_int, [_vp, _pcsr, _vp, C.POINTER(PcgOptions), _vp, _pint, _pdbl, _vp]),
    "ssn_components": (_int, [_vp, _pcsr, _vp, _vp, _vp, _vp, _pint]),
    "ssn_hybrid_amg": (_int, [_vp, C.POINTER(ProbData), C.POINTER(AmgOptions), _vp, _pint, _pdbl, _pint]),
    "ssn_hybrid_twogrid": (_int, [_vp, C.POINTER(ProbData), C.POINTER(AmgOptions), _vp, _pint, _pdbl, _pint]),
    "ssn_twogrid_bigph": (_int, [_vp, _pcsr, _vp, C.POINTER(AmgOptions), _vp, _pint, _pdbl, _vp, _vp, _pint]),
    "ssn_twogrid": (_int, [_vp, _pcsr, _vp, C.POINTER(AmgOptions), _vp, _pint, _pdbl, _vp, _vp, _pint]),
    "ssn_aug_pcg": (_int, [_vp, C.POINTER(ProbData), C.POINTER(PcgOptions), _vp, _pint, _pdbl, _pint]),
    "ssn_amg4pot": (_int, [_vp, C.POINTER(ProbData), C.POINTER(AmgOptions), _vp, _pint, _pdbl, _pint]),
    "ssn_pcg4pot": (_int, [_vp, C.POINTER(ProbData), C.POINTER(PcgOptions), _vp, _pint, _pdbl, _pint]),
    "ssn_rescaled_system": (_int, [_vp, C.POINTER(ProbData), _pcsr, _vp]),
    "ssn_jk_system": (_int, [_vp, C.POINTER(ProbData), _pcsr]),
    "ssn_spmv": (_int, [_vp, _pcsr, _vp, _vp]),
    "ssn_spgemm": (_int, [_vp, _pcsr, _pcsr, _pcsr]),
    "ssn_transpose": (_int, [_vp, _pcsr, _pcsr]),
}

_lib = None


def load():
    """dlopen the library and bind every declared symbol (no device needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` or `make -C {os.path.join(_HERE, 'csrc')}` -- there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Context:
    """One ``ssn_ctx`` on one CUDA device."""

    def __init__(self, device=-1):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.ssn_create(C.byref(h), int(device))
        if st != 0:
            raise SsnError(st, "ssn_create failed (no CUDA device? this package has no CPU fallback)")
        self.h = h

    def check(self, st):
        if st != 0:
            raise SsnError(st, self.lib.ssn_last_error(self.h).decode(errors="replace"))

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.h, *args))

    def launches(self):
        return int(self.lib.ssn_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.ssn_destroy(self.h)
            self.h = None


_ctx = {}


def context(device=None):
    import torch
    if not torch.cuda.is_available():
        raise SsnError(-1, "no CUDA device: the SsN-AMG operators run only on the GPU (no CPU fallback)")
    if device is None:
        device = torch.cuda.current_device()
    if device not in _ctx:
        with torch.cuda.device(device):
            _ctx[device] = Context(device)
    return _ctx[device]
