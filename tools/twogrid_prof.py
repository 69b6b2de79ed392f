"""Hybrid_twogrid (inner_solver = 5) on a saved SsN state: time per call and the phase profile (ncu target)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
from amg_prof import load_state  # noqa: E402


def main():
    path, tag = sys.argv[1], sys.argv[2]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    pd, m, n = load_state(path, tag)
    opts = ssnamg.driver.CLASS1_AMG_OPTIONS
    zs = []
    for rep in range(reps):
        if rep == reps - 1:
            ssnamg.profile(True)
        ssnamg.rng_reset(); l0 = ssnamg.launch_count(); torch.cuda.synchronize(); t0 = time.time()
        zeta, it, res, info = ssnamg.Hybrid_twogrid(pd, opts)
        torch.cuda.synchronize()
        print(f"{tag}: twogrid its={it} res={res:.2e} comps={info[0]} ms={(time.time() - t0) * 1e3:.2f} launches={ssnamg.launch_count() - l0}")
        zs.append(zeta.clone() if hasattr(zeta, "clone") else zeta)
    print(ssnamg.profile_dump())
    import ctypes
    from importlib import import_module
    lib = import_module("codes-of-ipd-ssn-amg-method_b200._lib")
    pb = (ctypes.c_ulonglong * 256)()
    lib.context().call("ssn_debug_cycles_persist", ctypes.cast(pb, ctypes.c_void_p), 1)
    ops = {0: "resid", 1: "gs_apply", 2: "jacobi", 3: "spmv(P)", 4: "leaf (dense / PCG)", 5: "outer res", 6: "zsum / PCG phases (8: A*p, 9: sum, 10: updates, 11: sum, 12: p + barrier)", 7: "kernel"}
    tot = pb[7 * 16] or 1
    for op, nm in ops.items():
        for lv in range(16):
            cnt = pb[128 + op * 16 + lv]
            if cnt:
                cyc = pb[op * 16 + lv]
                print(f"  pdbg {nm} slot {lv}: {cyc / 1e3:10.1f} kcycles ({100.0 * cyc / tot:5.1f} %) over {cnt:6d} calls -> {cyc / cnt:8.0f} cyc/call")
    if os.environ.get("SSN_TG_SAVE"):
        torch.save(zs[-1].cpu(), os.environ["SSN_TG_SAVE"])
    if os.environ.get("SSN_TG_COMPARE") and os.path.exists(os.environ["SSN_TG_COMPARE"]):
        ref = torch.load(os.environ["SSN_TG_COMPARE"])
        print("max rel diff against the saved solution:", float((zs[-1].cpu() - ref).abs().max() / ref.abs().max()))


if __name__ == "__main__":
    main()
