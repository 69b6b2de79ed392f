"""Phase profile of one SsN step of partial OT at the config-3 bench state (bench.py --config class2_64)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402

drv = ssnamg.driver
P = ssnamg.problems.grid_problem_pot(64, seed=0)
ssnamg.rng_reset()
st = drv.class2_trivial_state(P)
st["lk"], _, _ = drv.ssn_step_class2(st)
for _ in range(2):
    ssnamg.rng_reset(); drv.ssn_step_class2(st)
ssnamg.profile(True)
for _ in range(3):
    ssnamg.rng_reset(); _, _, info = drv.ssn_step_class2(st)
torch.cuda.synchronize()
print(ssnamg.profile_dump())
print({k: info[k] for k in ("E", "nnzH", "itamg", "ll", "ms_plan", "ms_asat", "ms_amg")})
