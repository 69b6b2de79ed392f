"""Runs the device-resident Class1 solve on a grid problem and prints per-step statistics."""
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
    g = int(sys.argv[1]); max_s = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    t0 = time.time()
    P = ssnamg.problems.grid_problem(g, seed=0)
    print(f"g={g} m=n={g * g} gen {time.time() - t0:.1f}s", flush=True)
    recs = []

    def hook(st):
        recs.append((st["k"], st["ssn_it"], st["E"], st["H0"].nnz))
    ssnamg.rng_reset()
    l0 = ssnamg.launch_count()
    out = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], on_ssn_step=hook, verbose=True, max_seconds=max_s)
    st = out["stats"]
    print(f"RESULT g={g} converged={st['converged']} outer={out['outer_its']} relKKT={out['rel_kkt']:.3e} f={out['fxk'][-1]:.10f} "
          f"ssn={sum(st['ssn_its'])} ls_trials={st['ls_trials']} amg_calls={st['amg_calls']} loop_s={out['seconds']:.2f} "
          f"warmup_s={out['warmup_seconds']:.2f} solve_s={st['solve_s']:.2f} asat_s={st['asat_s']:.2f} plan_s={st['plan_s']:.2f} "
          f"launches={ssnamg.launch_count() - l0}")
    calls = np.array([(e, t) for e, t, _, _ in st["solve_calls"]]) if st["solve_calls"] else np.zeros((0, 2))
    for lo, hi in ((0, 1), (1, 1e5), (1e5, 1e6), (1e6, 4e6), (4e6, 1e9)):
        sel = (calls[:, 0] >= lo) & (calls[:, 0] < hi)
        if sel.any():
            print(f"  AMG calls with E in [{lo:g},{hi:g}): {int(sel.sum()):4d} calls, {calls[sel, 1].sum():7.2f} s, mean {1e3 * calls[sel, 1].mean():7.1f} ms, max {1e3 * calls[sel, 1].max():7.1f} ms")
    E = np.array([r[2] for r in recs])
    if E.size:
        print("E min/median/max", E.min(), np.median(E), E.max(), " lin its", [i for its in st["lin_its"] for i in its][:60])


if __name__ == "__main__":
    main()
