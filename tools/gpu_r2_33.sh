# round 2, call 33: pull form of the cluster-wide sum (z_sum1_pull) -- micro-benchmark and A/B of Hybrid_AMG at the benchmarked state
cd $GRAFT_REPO_ROOT
timeout 300 python tools/barrier_bench.py 2>&1 | grep -E "z_sum1|z_barrier|block reduction" 
for tag in "" "_pull" "" "_pull"; do
  echo "== libssnamg$tag.so"
  SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg$tag.so timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 2>&1 | grep -E "k30_s1" | tail -3
done
SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg_pull.so timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -x > gpurun_out/pytest_amg_pull.log 2>&1; echo "pytest(pull) rc=$?"; tail -2 gpurun_out/pytest_amg_pull.log
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q -x -k "class2" > gpurun_out/pytest_r2ae.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r2ae.log
