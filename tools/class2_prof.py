"""Phase profile of one SsN step of partial OT at the config-3 bench state (bench.py --config class2_64)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402

drv = ssnamg.driver
P = ssnamg.problems.grid_problem_pot(64, seed=0)
ssnamg.rng_reset()
st = drv.class2_trivial_state(P)
st["lk"], _, _ = drv.ssn_step_class2(st)
for _ in range(2):
    ssnamg.rng_reset(); drv.ssn_step_class2(st)
ssnamg.profile(True)
for _ in range(3):
    ssnamg.rng_reset(); _, _, info = drv.ssn_step_class2(st)
torch.cuda.synchronize()
print(ssnamg.profile_dump())
print({k: info[k] for k in ("E", "nnzH", "itamg", "ll", "ms_plan", "ms_asat", "ms_amg")})
# per-(operation, level) cycles of the grid-wide solve kernel (debug build only: SSN_LIB_PATH=.../libssnamg_dbg.so)
import ctypes
from importlib import import_module
_lib = import_module("codes-of-ipd-ssn-amg-method_b200._lib")
_pb = (ctypes.c_ulonglong * 256)()
_lib.context().call("ssn_debug_cycles_persist", ctypes.cast(_pb, ctypes.c_void_p), 1)
_ops = {0: "resid", 1: "gs_apply", 2: "jacobi", 3: "spmv(P)", 4: "dense", 5: "outer res", 6: "zsum/dots", 7: "kernel"}
_tot = _pb[7 * 16] or 1
for _op, _nm in _ops.items():
    for _lv in range(16):
        _cnt = _pb[128 + _op * 16 + _lv]
        if _cnt:
            _cyc = _pb[_op * 16 + _lv]
            print(f"  pdbg {_nm:9s} level {_lv}: {_cyc / 1e3:10.1f} kcycles ({100.0 * _cyc / _tot:5.1f} %) over {_cnt:6d} calls -> {_cyc / _cnt:8.0f} cyc/call")
