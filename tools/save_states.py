"""Runs the device Class1 solve on a g x g grid problem and saves compact SsN states (active-set
coordinates, right-hand side, bk1, tk) so that later profiling runs need not repeat the solve."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
drv = ssnamg.driver


def main():
    g = int(sys.argv[1]); picks = [tuple(int(v) for v in t.split(".")) for t in sys.argv[2].split(",")]
    out = sys.argv[3]
    P = ssnamg.problems.grid_problem(g, seed=0)
    rec = {}

    def hook(st):
        key = (st["k"], st["ssn_it"])
        if key in picks:
            lin = torch.nonzero(st["s"]).reshape(-1).cpu().numpy().astype(np.int64)
            tag = f"k{key[0]}_s{key[1]}"
            rec[tag + "_lin"] = lin; rec[tag + "_z"] = (-st["Fk"]).cpu().numpy()
            rec[tag + "_bk1"] = st["bk1"]; rec[tag + "_tk"] = st["tk"]
            print("saved", tag, "E", lin.size, flush=True)
    ssnamg.rng_reset()
    drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], on_ssn_step=hook, max_outer=max(k for k, _ in picks))
    rec["g"] = g
    np.savez_compressed(out, **rec)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
