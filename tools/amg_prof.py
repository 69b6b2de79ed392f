"""Hybrid_AMG on a saved SsN state (tools/save_states.py): timing + phase profile; ncu target."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def load_state(path, tag):
    d = np.load(path)
    g = int(d["g"]); m = n = g * g
    s = torch.zeros(m * n, dtype=torch.uint8, device="cuda")
    s[torch.from_numpy(d[tag + "_lin"]).cuda()] = 1
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    H = ssnamg.ASAt(s, p, q)
    del s
    return {"bk1": float(d[tag + "_bk1"]), "tk": float(d[tag + "_tk"]), "p": p, "q": q, "T": None, "H0": H,
            "z": torch.from_numpy(d[tag + "_z"]).cuda()}, m, n


def main():
    path, tag = sys.argv[1], sys.argv[2]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    prof = len(sys.argv) > 4 and sys.argv[4] == "prof"
    pd, m, n = load_state(path, tag)
    opts = ssnamg.driver.CLASS1_AMG_OPTIONS
    ssnamg.profile(prof)
    for rep in range(reps):
        ssnamg.rng_reset(); l0 = ssnamg.launch_count(); torch.cuda.synchronize(); t0 = time.time()
        zeta, it, res, info = ssnamg.Hybrid_AMG(pd, opts)
        torch.cuda.synchronize()
        print(f"{tag}: nnz(H0)={pd['H0'].nnz} comps={info[0]} cycles={it} res={res:.1e} ms={(time.time() - t0) * 1e3:.2f} launches={ssnamg.launch_count() - l0}")
    if prof:
        print(ssnamg.profile_dump())
    import ctypes
    from importlib import import_module
    lib = import_module("codes-of-ipd-ssn-amg-method_b200._lib")
    ctx = lib.context(); buf = (ctypes.c_ulonglong * 64)()
    ctx.call("ssn_debug_cycles", ctypes.cast(buf, ctypes.c_void_p), 1)
    pb = (ctypes.c_ulonglong * 256)()
    ctx.call("ssn_debug_cycles_persist", ctypes.cast(pb, ctypes.c_void_p), 1)
    ops = {0: "resid", 1: "gs_apply", 2: "jacobi", 3: "spmv(P)", 4: "dense", 5: "outer res", 6: "zsum/dots", 7: "kernel"}
    tot = pb[7 * 16] or 1
    for op, nm in ops.items():
        for lv in range(16):
            cnt = pb[128 + op * 16 + lv]
            if cnt:
                cyc = pb[op * 16 + lv]
                print(f"  pdbg {nm:9s} level {lv}: {cyc / 1e3:10.1f} kcycles ({100.0 * cyc / tot:5.1f} %) over {cnt:6d} calls -> {cyc / cnt:8.0f} cyc/call")
    names = {0: "smooth", 8: "resid+restrict", 16: "prolong", 24: "pcg"}
    for base, nm in names.items():
        for k in range(8 if base < 24 else 1):
            if buf[32 + base + k]:
                print(f"  dbg {nm:15s} level+{k}: {buf[base + k] / 1e3:10.1f} kcycles over {buf[32 + base + k]} calls -> {buf[base + k] / buf[32 + base + k]:9.0f} cyc/call")


if __name__ == "__main__":
    main()
