import os, sys, time, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import ssnamg
drv = ssnamg.driver
P = ssnamg.problems.grid_problem_pot(64, seed=0)
ssnamg.rng_reset()
st = drv.class2_trivial_state(P)
st["lk"], _, _ = drv.ssn_step_class2(st)
out = []
for i in range(24):
    ssnamg.rng_reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
    _, _, info = drv.ssn_step_class2_ops(st)
    torch.cuda.synchronize(); out.append((round((time.perf_counter() - t0) * 1e3, 1), round(info["ms_amg"], 1)))
print("ops path (total ms, amg ms):", out)
out = []
for i in range(24):
    ssnamg.rng_reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
    drv.ssn_step_class2(st)
    torch.cuda.synchronize(); out.append(round((time.perf_counter() - t0) * 1e3, 1))
print("one-call path ms:", out)
