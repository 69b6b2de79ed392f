"""Generates tests/golden/trace_*.npz: the CPU oracle's first outer iterations (or whole solve) on
BASELINE.json's own configurations, recorded step by step, for the `-m gpu` tests that pin the
device-resident drivers to them (tests/test_gpu_traces.py).

  python tests/golden/make_traces.py class1_grid64 [outer]     config 2: 64x64 grids, Class 1
  python tests/golden/make_traces.py class2_grid64 [outer]     config 3: 64x64 grids, Class 2 (partial OT)
  python tests/golden/make_traces.py bundled500                config 1: the reference's own input, whole solve
                                                               (+ bundled500_inputs.npz: c, r, l of the .mat file,
                                                               which does not exist on the GPU box)
  python tests/golden/make_traces.py class2_bundled500         config 3's small fixture: data4-500.mat, whole solve

Recorded per outer iteration: objective fxk, KKT residuals; per SsN step: (k, ssn_it, E = nnz(s), components,
inner iterations, accepted backtracking exponent ll, |F| after the step); the final duals lk; of the final
plan: nnz, sum, 2-norm, inf-norm and its values at the 4096 largest entries + 4096 fixed random positions.
MATLAB/Octave are absent, so these are the oracle's numbers (parity unpinned w.r.t. MATLAB).
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                                     # noqa: E402
from oracle import driver                         # noqa: E402
problems = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.problems")
OUT = os.path.dirname(os.path.abspath(__file__))


def plan_digest(x, seed=1234):
    top = np.argsort(-x, kind="stable")[:4096].astype(np.int64)
    rnd = np.random.RandomState(seed).randint(0, x.size, 4096).astype(np.int64)
    idx = np.concatenate([top, rnd])
    return {"x_nnz": int(np.count_nonzero(x)), "x_sum": float(x.sum()), "x_norm2": float(np.linalg.norm(x)),
            "x_inf": float(np.abs(x).max()), "x_idx": idx, "x_val": x[idx]}


def record(tag, out, extra=None):
    rec = {"fxk": np.array(out["fxk"]), "outer_its": out["outer_its"], "rel_kkt": out["rel_kkt"],
           "converged": bool(out["stats"]["converged"]), "ssn_its": np.array(out["stats"]["ssn_its"]),
           "steps": np.array(out["stats"]["steps"], dtype=np.float64), "lk": out["lk"], "seconds": out["seconds"]}
    if "KKT" in out:
        rec["KKT"] = np.array(out["KKT"])
    else:
        rec["KKT_xk"] = np.array(out["KKT_xk"]); rec["KKT_lk"] = np.array(out["KKT_lk"])
    rec.update(plan_digest(out["xk"]))
    rec.update(extra or {})
    path = os.path.join(OUT, f"trace_{tag}.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "outer", out["outer_its"], "relKKT", out["rel_kkt"], "f", out["fxk"][-1],
          "ssn", rec["ssn_its"].tolist(), "seconds", round(out["seconds"], 1), flush=True)


def main():
    what = sys.argv[1]
    outer = int(sys.argv[2]) if len(sys.argv) > 2 else None
    t0 = time.time()
    oracle.rng_reset()
    if what == "class1_grid64":
        P = problems.grid_problem(64, seed=0)
        out = driver.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], max_outer=outer or 4, verbose=True)
        record(f"class1_grid64_outer{outer or 4}", out, {"g": 64})
    elif what == "class2_grid64":
        P = problems.grid_problem_pot(64, seed=0)
        out = driver.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], max_outer=outer or 3, verbose=True)
        record(f"class2_grid64_outer{outer or 3}", out, {"g": 64, "mu": P["mu"], "mass": float(P["phi"] @ out["xk"])})
    elif what == "class2_grid64_nowarm":
        # the APD / SsN loop of Class 2 from the trivial start (no A-ADMM iterations): the closed-form inverse of the warm
        # start (Class2/invHHt.m:8-9, s = t - l'*Vl with t ~ m*n: "not robust", invAAt.m:6) loses ~eps*m*n, so two
        # implementations leave the 100-iteration warm start of a 64x64 problem ~1e-3 apart in the duals and the loop
        # cannot be compared step by step from there; from a common start it can
        P = problems.grid_problem_pot(64, seed=0)
        out = driver.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], max_outer=outer or 3, warm_maxit=0, verbose=True)
        record(f"class2_grid64_nowarm_outer{outer or 3}", out, {"g": 64, "mu": P["mu"], "mass": float(P["phi"] @ out["xk"])})
    elif what == "class1_grid":
        g = int(sys.argv[2]); outer = int(sys.argv[3]) if len(sys.argv) > 3 else None
        P = problems.grid_problem(g, seed=0)
        out = driver.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], max_outer=outer, verbose=True)
        record(f"class1_grid{g}" + (f"_outer{outer}" if outer else ""), out, {"g": g})
    elif what == "class2_grid":
        g = int(sys.argv[2]); outer = int(sys.argv[3]) if len(sys.argv) > 3 else None
        P = problems.grid_problem_pot(g, seed=0)
        out = driver.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], max_outer=outer, verbose=True)
        record(f"class2_grid{g}" + (f"_outer{outer}" if outer else ""), out, {"g": g, "mu": P["mu"], "mass": float(P["phi"] @ out["xk"])})
    elif what == "bundled500":
        P = problems.load_bundled_class1("/root/reference/Class1/InputData/data1-500.mat")
        np.savez_compressed(os.path.join(OUT, "bundled500_inputs.npz"), c=P["c"], r=P["r"], l=P["l"], m=P["m"], n=P["n"])
        out = driver.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], verbose=True)
        record("bundled500", out)
    elif what == "class2_bundled500":
        P = problems.load_bundled_class2("/root/reference/Class2/InputData/data4-500.mat")
        np.savez_compressed(os.path.join(OUT, "bundled500_class2_inputs.npz"), c=P["c"], r=P["r"], l=P["l"], mu=P["mu"],
                            m=P["m"], n=P["n"])
        out = driver.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], verbose=True)
        record("class2_bundled500", out, {"mu": P["mu"], "mass": float(P["phi"] @ out["xk"])})
    else:
        raise SystemExit(__doc__)
    print("total", round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    main()
