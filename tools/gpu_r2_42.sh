# round 2, call 42: batched tail of z_row_dot (and the gather table through the read-only path) at the benchmarked state
cd $GRAFT_REPO_ROOT
for tag in "" "_nc" "" "_nc"; do
  echo "== libssnamg$tag.so"
  SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg$tag.so timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 prof 2>&1 | grep -E "k30_s1|solve.dsm_solve_kernel  " | tail -3
done
echo "== halo on"; SSN_DSM_HALO=1 timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 prof 2>&1 | grep -E "k30_s1|solve.dsm_solve_kernel  " | tail -2
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -x > gpurun_out/pytest_amg_r2ak.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_amg_r2ak.log
SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg_nc.so timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -x > gpurun_out/pytest_amg_nc_r2ak.log 2>&1; echo "pytest(nc) rc=$?"; tail -2 gpurun_out/pytest_amg_nc_r2ak.log
