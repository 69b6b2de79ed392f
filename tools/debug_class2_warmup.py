"""Device vs oracle warm start of Class 2 (Class2/warmup_class2.m) at growing sizes / iteration counts."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
from oracle import driver as odrv  # noqa: E402

drv = ssnamg.driver
for g in (8, 16, 32):
    P = ssnamg.problems.grid_problem_pot(g, seed=0)
    for it in (1, 2, 5, 20, 100):
        u_ref, l_ref = odrv.warmup_class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], 0, it)
        u, l = drv.warmup_class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], 0.0, it)
        u = u.cpu().numpy(); l = l.cpu().numpy()
        print(f"g={g:3d} it={it:3d}  |u-u_ref|/|u_ref| = {np.linalg.norm(u - u_ref) / max(np.linalg.norm(u_ref), 1e-300):.2e}   "
              f"|l-l_ref|/|l_ref| = {np.linalg.norm(l - l_ref) / max(np.linalg.norm(l_ref), 1e-300):.2e}   max|l| {np.abs(l_ref).max():.3e}", flush=True)
