"""Runs the device-resident Class2 (partial OT) solve on a g x g grid problem (config 3 of BASELINE.json)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
drv = ssnamg.driver


def main():
    g = int(sys.argv[1]); max_s = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
    solver = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    P = ssnamg.problems.grid_problem_pot(g, seed=0)
    ssnamg.rng_reset(); l0 = ssnamg.launch_count(); t0 = time.time()
    out = drv.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], inner_solver=solver, max_seconds=max_s,
                             verbose="-v" in sys.argv)
    st = out["stats"]
    x = out["xk"]
    print(f"RESULT class2 g={g} solver={solver} converged={st['converged']} outer={out['outer_its']} relKKT={out['rel_kkt']:.3e} "
          f"f={out['fxk'][-1]:.10f} mass={float(x.sum()):.8f} mu={P['mu']:.8f} ssn={sum(st['ssn_its'])} ls_trials={st['ls_trials']} "
          f"amg_calls={st['amg_calls']} loop_s={out['seconds']:.2f} warmup_s={out['warmup_seconds']:.2f} total_s={time.time() - t0:.2f} "
          f"launches={ssnamg.launch_count() - l0}")


if __name__ == "__main__":
    main()
