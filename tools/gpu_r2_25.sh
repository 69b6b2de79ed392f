# round 2, call 25: shared-memory copy of the gathered vector on dense levels (grid-wide kernel): A/B on the class 2 step, tests
cd $GRAFT_REPO_ROOT
for v in 0 1; do
  SSN_STAGE_DENSE=$v timeout 600 python tools/class2_prof.py > gpurun_out/class2_prof_stage${v}_r2x.log 2>&1; echo "stage_dense=$v rc=$?"
  grep -E "persist_solve_kernel  |amg_setup total  |ms_amg" gpurun_out/class2_prof_stage${v}_r2x.log | tail -3
done
timeout 1500 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py tests/test_gpu_traces.py -m gpu -q > gpurun_out/pytest_amg_r2x.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_amg_r2x.log | tail -6
