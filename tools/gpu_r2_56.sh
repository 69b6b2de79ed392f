# round 2, call 56: PCG leaf with a row's d, r, own p and diagonal in registers: timing at the two saved states, the two-grid tests
cd $GRAFT_REPO_ROOT
for st in k30_s1 k80_s2; do timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz $st 5 2>&1 | grep -E "^$st|solve.dsm_solve_kernel  " | tail -3; done
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_driver.py tests/test_gpu_traces.py -m gpu -q -x -s 2>&1 | grep -E "two-grid in the cluster|passed|failed|rror" | tail -8
