"""Phase profile of Hybrid_AMG on realistic SsN states captured from the device driver."""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
drv = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.driver")


def main():
    g = int(sys.argv[1]); outer = [int(v) for v in sys.argv[2].split(",")]
    P = ssnamg.problems.grid_problem(g, seed=0)
    states = {}

    def hook(st):
        if st["k"] in outer and st["ssn_it"] == 2:
            states[st["k"]] = {"s": st["s"].clone(), "z": (-st["Fk"]).clone(), "bk1": st["bk1"], "tk": st["tk"], "E": st["E"]}
    ssnamg.rng_reset()
    drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], on_ssn_step=hook, max_outer=max(outer))
    m = n = g * g
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    for k, st in sorted(states.items()):
        H = ssnamg.ASAt(st["s"], p, q)
        pd = {"bk1": st["bk1"], "tk": st["tk"], "p": p, "q": q, "T": None, "H0": H, "z": st["z"]}
        for prof in (False, True):
            ssnamg.profile(prof)
            ts = []
            for rep in range(3):
                ssnamg.rng_reset(); torch.cuda.synchronize(); t0 = time.time()
                zeta, it, res, info = ssnamg.Hybrid_AMG(pd, drv.CLASS1_AMG_OPTIONS)
                torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
            print(f"outer {k}: E={st['E']} nnz(H0)={H.nnz} comps={info[0]} cycles={it} res={res:.1e} prof={prof} ms={['%.2f' % t for t in ts]}")
            if prof:
                print(ssnamg.profile_dump())
        ssnamg.profile(False)


if __name__ == "__main__":
    main()
