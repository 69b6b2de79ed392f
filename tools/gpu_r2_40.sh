# round 2, call 40: halo copies of the A-gathers in dsm_solve_kernel (SSN_DSM_HALO) A/B at the benchmarked state; AMG tests
cd $GRAFT_REPO_ROOT
for h in 1 0 1 0; do echo "== SSN_DSM_HALO=$h"; SSN_DSM_HALO=$h timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 prof 2>&1 | grep -E "k30_s1|halo entries|solve.dsm_solve_kernel  " | tail -5; done
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -x > gpurun_out/pytest_amg_r2aj.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_amg_r2aj.log
for tag in k12_s2 k40_s2 k80_s2; do SSN_DSM_HALO=1 timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz $tag 4 prof 2>&1 | grep -E "$tag|halo entries" | tail -2; SSN_DSM_HALO=0 timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz $tag 4 2>&1 | grep -E "$tag" | tail -1; done
