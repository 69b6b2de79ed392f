"""Prints the device driver's per-step record next to the oracle's recorded trace (tests/golden/trace_*.npz):
    python tools/trace_compare.py class2_grid64_outer3 | class1_grid64_outer4 | bundled500 | class2_bundled500 [native]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssnamg  # noqa: E402


def main():
    name = sys.argv[1]; native = len(sys.argv) > 2 and sys.argv[2] == "native"
    T = dict(np.load(os.path.join(ROOT, "tests", "golden", f"trace_{name}.npz")))
    drv = ssnamg.driver
    k = int(T["outer_its"])
    ssnamg.rng_reset()
    if name.startswith("class2_grid"):
        P = ssnamg.problems.grid_problem_pot(int(T["g"]), seed=0)
        out = drv.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], max_outer=k, verbose=True)
    elif name.startswith("class1_grid"):
        P = ssnamg.problems.grid_problem(int(T["g"]), seed=0)
        f = ssnamg.APD_SsN_Class1 if native else drv.APD_SsN_Class1
        out = f(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], max_outer=k, verbose=True)
    elif name == "bundled500":
        D = np.load(os.path.join(ROOT, "tests", "golden", "bundled500_inputs.npz")); m, n = int(D["m"]), int(D["n"])
        f = ssnamg.APD_SsN_Class1 if native else drv.APD_SsN_Class1
        out = f(D["c"], D["r"], D["l"], np.ones(m), np.ones(n), np.inf)
    else:
        D = np.load(os.path.join(ROOT, "tests", "golden", "bundled500_class2_inputs.npz")); m, n = int(D["m"]), int(D["n"])
        out = drv.APD_SsN_Class2(D["c"], D["r"], D["l"], np.ones(m), np.ones(n), float(D["mu"]), np.ones(m * n))
    print("fxk   dev", out["fxk"][:6]); print("fxk   ref", T["fxk"][:6].tolist())
    if "KKT" in T:
        print("KKT0  dev", out["KKT"][0]); print("KKT0  ref", T["KKT"][0].tolist())
    else:
        print("KKT0  dev", out["KKT_xk"][0], out["KKT_lk"][0]); print("KKT0  ref", float(T["KKT_xk"][0]), float(T["KKT_lk"][0]))
    print("ssn_its dev", out["stats"]["ssn_its"]); print("ssn_its ref", T["ssn_its"].tolist())
    g = np.array(out["stats"]["steps"], dtype=np.float64).reshape(-1, 7); r = T["steps"].reshape(-1, 7)
    print("   k  it |        E dev        E ref | comp dev ref | its dev ref |  ll dev ref |      |F| dev      |F| ref")
    for i in range(max(len(g), len(r))):
        a = g[i] if i < len(g) else [np.nan] * 7; b = r[i] if i < len(r) else [np.nan] * 7
        flag = "" if (i < len(g) and i < len(r) and np.array_equal(a[:6], b[:6])) else "   <--"
        print(f"{a[0]:4.0f} {a[1]:3.0f} | {a[2]:12.0f} {b[2]:12.0f} | {a[3]:6.0f} {b[3]:6.0f} | {a[4]:4.0f} {b[4]:4.0f} | {a[5]:4.0f} {b[5]:4.0f} | {a[6]:12.5e} {b[6]:12.5e}{flag}")


if __name__ == "__main__":
    main()
