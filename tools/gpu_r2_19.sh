# round 2, call 19: one-launch MIS rounds: A/B, AMG tests
cd $GRAFT_REPO_ROOT
for v in 0 1; do
  SSN_MIS_CLUSTER=$v timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 6 prof > gpurun_out/amg_prof_mis${v}_r2s.log 2>&1; echo "amg_prof mis_cluster=$v rc=$?"
  grep -E "k30_s1|amg_setup total|setup.mis_set|rror" gpurun_out/amg_prof_mis${v}_r2s.log | tail -10
done
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q > gpurun_out/pytest_amg_r2s.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_amg_r2s.log | tail -6
