"""Phase profile of Hybrid_AMG at the benchmarked SsN state (outer 30, SsN step 1 of the grid solve):
setup phases, launch counts, persistent solve kernel.  Development aid / ncu target.
Usage: python tools/amg_state_prof.py [grid=128] [outer=30]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    outer = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    drv = ssnamg.driver
    P = ssnamg.problems.grid_problem(g, seed=0)
    ssnamg.rng_reset()
    st = drv.capture_state(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], outer=outer, ssn_it=1)
    del P
    ev = ssnamg.prox_residual(st["wk"], st["lk"], st["p"], st["q"], st["tk"], float("inf"), want=("Axprox", "s"))
    H0 = ssnamg.ASAt(ev["s"], st["p"], st["q"])
    Fk = st["bk1"] * st["lk"] - ev["Axprox"] - st["wlk"]
    pd = {"bk1": st["bk1"], "tk": st["tk"], "q": st["q"], "p": st["p"], "T": None, "H0": H0, "z": -Fk}
    opts = drv.CLASS1_AMG_OPTIONS
    for prof in (False, True):
        ssnamg.profile(prof)
        for rep in range(3):
            ssnamg.rng_reset(); l0 = ssnamg.launch_count(); torch.cuda.synchronize(); t0 = time.perf_counter()
            zeta, it, res, info = ssnamg.Hybrid_AMG(pd, opts)
            torch.cuda.synchronize()
            print(f"prof={prof} nnz(H0)={H0.nnz} comps={info[0]} cycles={it} res={res:.1e} ms={(time.perf_counter() - t0) * 1e3:.2f} "
                  f"launches={ssnamg.launch_count() - l0}")
        if prof:
            print(ssnamg.profile_dump())
    ssnamg.profile(False)
    for rep in range(3):                                    # the two-level alternative (inner_solver = 5) on the same system
        ssnamg.rng_reset(); l0 = ssnamg.launch_count(); torch.cuda.synchronize(); t0 = time.perf_counter()
        zeta2, it2, res2, info2 = ssnamg.Hybrid_twogrid(pd, opts)
        torch.cuda.synchronize()
        print(f"Hybrid_twogrid: comps={info2[0]} its={it2} res={res2:.1e} ms={(time.perf_counter() - t0) * 1e3:.2f} "
              f"launches={ssnamg.launch_count() - l0} |zeta - zeta_amg|/|zeta_amg|={float(torch.linalg.norm(zeta2 - zeta) / torch.linalg.norm(zeta)):.1e}")


if __name__ == "__main__":
    main()
