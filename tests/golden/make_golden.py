"""Generates tests/golden/*.npz from the CPU oracle (and, for the bundled-data systems, from the
reference's own input file Class1/InputData/data1-500.mat, read here in the build container
only).  Run from the repository root:  python tests/golden/make_golden.py

Each `ssn_system_*.npz` holds one semismooth-Newton linear system of the reference's default
configuration (inner_solver = 4) together with what the oracle computed for it:
  inputs : m, n, s (packed bits, column-major), z (= -Fk_old), bk1, tk
  outputs: H0 (CSC arrays), level sizes, per-level C/F vectors, W-cycle count, residual history,
           zeta, number of components, random numbers drawn.
"""
import importlib
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                                     # noqa: E402
from oracle import driver, amg                    # noqa: E402
problems = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.problems")
OUT = os.path.dirname(os.path.abspath(__file__))


def solve_and_record(tag, P, picks, max_seconds=None):
    snaps = []

    def hook(st):
        idx = len(hook.seen); hook.seen.append(1)
        if idx in picks:
            snaps.append({"idx": idx, "k": st["k"], "ssn_it": st["ssn_it"], "s": st["s"].copy(),
                          "z": -st["Fk_old"].copy(), "bk1": st["bk1"], "tk": st["tk"]})
    hook.seen = []
    oracle.rng_reset()
    out = driver.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], on_ssn_step=hook,
                                max_seconds=max_seconds)
    print(tag, "outer", out["outer_its"], "relKKT", out["rel_kkt"], "f", out["fxk"][-1], "systems", len(hook.seen))
    m, n = P["m"], P["n"]
    for sn in snaps:
        H0 = oracle.ASAt(sn["s"], P["p"], P["q"])
        pd = {"bk1": sn["bk1"], "tk": sn["tk"], "p": P["p"], "q": P["q"], "T": sp.diags(np.zeros(m + n)),
              "H0": H0, "z": sn["z"]}
        oracle.rng_reset()
        zeta, it, res, info = oracle.Hybrid_AMG(pd, driver.CLASS1_AMG_OPTIONS)
        drawn = oracle.GLOBAL_STREAM.drawn
        last = amg.amg_state.last
        rec = {"m": m, "n": n, "s_bits": np.packbits(sn["s"].astype(np.uint8)), "z": sn["z"], "bk1": sn["bk1"],
               "tk": sn["tk"], "H_indptr": H0.indptr.astype(np.int32), "H_indices": H0.indices.astype(np.int32),
               "H_data": H0.data, "zeta": zeta, "it": it, "res": res, "info": np.asarray(info), "drawn": drawn,
               "outer_k": sn["k"], "ssn_it": sn["ssn_it"]}
        if info[0] == 1 and last is not None:
            rec["level_sizes"] = np.array([A.shape[0] for A in last["Ack"]])
            rec["level_nnz"] = np.array([A.nnz for A in last["Ack"]])
            for lvl, tr in enumerate(last["trace"]):
                rec[f"isC_{lvl + 2}"] = np.packbits(tr["isC"].astype(np.uint8))
        np.savez_compressed(os.path.join(OUT, f"ssn_system_{tag}_{sn['idx']:03d}.npz"), **rec)
        print("  saved", tag, sn["idx"], "E", int(sn["s"].sum()), "info", info, "it", it, "res", res,
              "levels", rec.get("level_sizes"))
    return out


def main():
    ref = "/root/reference/Class1/InputData/data1-500.mat"
    if os.path.exists(ref):
        P = problems.load_bundled_class1(ref)
        out = solve_and_record("bundled500", P, picks={2, 5, 40, 90, 140, 150})
        np.savez_compressed(os.path.join(OUT, "bundled500_summary.npz"), outer_its=out["outer_its"],
                            rel_kkt=out["rel_kkt"], f=out["fxk"][-1], nnz=np.count_nonzero(out["xk"]),
                            ssn_its=np.array(out["stats"]["ssn_its"]))
    P = problems.grid_problem(12, seed=0)       # 144 x 144 grid problem, full solve
    out = solve_and_record("grid12", P, picks={1, 4, 10, 20})
    np.savez_compressed(os.path.join(OUT, "grid12_summary.npz"), outer_its=out["outer_its"], rel_kkt=out["rel_kkt"],
                        f=out["fxk"][-1], ssn_its=np.array(out["stats"]["ssn_its"]))
    # MATLAB's well-known first draws of the start-up stream (rand after a fresh start)
    oracle.rng_reset()
    np.savez(os.path.join(OUT, "matlab_rand_first.npz"), first=oracle.rand(1000))


if __name__ == "__main__":
    main()


def bundled_class2_summary():
    """Summary of the oracle's Class2 solve on the reference's bundled Class2/InputData/data4-500.mat
    (objective, iteration counts, transported mass) -> bundled500_class2_summary.npz (about 35 s)."""
    import scipy.io
    from oracle import driver as odrv
    d = scipy.io.loadmat("/root/reference/Class2/InputData/data4-500.mat", mat_dtype=True)
    f = lambda k: np.ascontiguousarray(d[k], dtype=np.float64).reshape(-1)
    c, r, l, p, q, phi = f("c"), f("r"), f("l"), f("p"), f("q"), f("phi"); mu = float(d["mu"].reshape(-1)[0])
    oracle.rng_reset()
    out = odrv.APD_SsN_Class2(c, r, l, p, q, mu, phi)
    np.savez_compressed(os.path.join(OUT, "bundled500_class2_summary.npz"), outer_its=out["outer_its"], rel_kkt=out["rel_kkt"],
                        fxk=np.array(out["fxk"]), ssn_its=np.array(out["stats"]["ssn_its"]),
                        nnz_x=int((out["xk"] > 1e-9).sum()), mass=float(phi @ out["xk"]), mu=mu)


if __name__ == "__main__" and "class2" in sys.argv:
    bundled_class2_summary()

