# round 2, call 5: fused small-level setup + cluster/grid solve kernels with shared-memory control state; class 2 trace debugging
cd $GRAFT_REPO_ROOT
D=$PWD/codes-of-ipd-ssn-amg-method_b200
SSN_LIB_PATH=$D/libssnamg_dbg.so timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 prof > gpurun_out/amg_prof_cluster_dbg_r2e.log 2>&1; echo "amg_prof dbg cluster rc=$?"
grep -E "k30_s1|pdbg|cluster_solve_kernel  |amg_setup total  |fused_small" gpurun_out/amg_prof_cluster_dbg_r2e.log | head -24
SSN_CLUSTER_SOLVE=0 SSN_LIB_PATH=$D/libssnamg_dbg.so timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 prof > gpurun_out/amg_prof_grid_dbg_r2e.log 2>&1; echo "amg_prof dbg grid rc=$?"
grep -E "k30_s1|pdbg|persist_solve_kernel  " gpurun_out/amg_prof_grid_dbg_r2e.log | head -20
for v in "SSN_FUSED_SETUP=0" "SSN_FUSED_SETUP=1"; do
  env $v timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_release_${v}_r2e.log 2>&1; echo "amg_prof release $v rc=$?"
  grep -E "k30_s1|cluster_solve_kernel  |amg_setup total  |fused_small|transfer\(mis|build_dense" gpurun_out/amg_prof_release_${v}_r2e.log | tail -9
done
timeout 600 python tools/trace_compare.py class2_grid64_outer3 > gpurun_out/trace_c2g64.log 2>&1; echo "trace c2g64 rc=$?"; grep -v "^APD\|^   SsN" gpurun_out/trace_c2g64.log | head -50
timeout 1500 python -m pytest tests/test_abi.py tests/test_gpu_amg.py tests/test_gpu_solvers.py tests/test_gpu_traces.py -m gpu -q -s > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|^Class_AMG|passed|failed|error|Error" gpurun_out/pytest_gpu_r2e.log | tail -30
