# round 2, call 4: cluster kernel variants (L1-cached gathers, U rows in flight), parity suites that touch the AMG path
cd $GRAFT_REPO_ROOT
D=$PWD/codes-of-ipd-ssn-amg-method_b200
for tag in "" "_t512"; do
  SSN_LIB_PATH=$D/libssnamg_dbg$tag.so timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 prof > gpurun_out/amg_prof_cluster_dbg${tag}_r2d.log 2>&1; echo "amg_prof dbg$tag rc=$?"
  grep -E "k30_s1|pdbg|cluster_solve_kernel  " gpurun_out/amg_prof_cluster_dbg${tag}_r2d.log | head -20
done
timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_release_r2d.log 2>&1; echo "amg_prof release rc=$?"
grep -E "k30_s1|cluster_solve_kernel  |amg_setup total  " gpurun_out/amg_prof_release_r2d.log | head
timeout 1500 python -m pytest tests/test_abi.py tests/test_gpu_amg.py tests/test_gpu_solvers.py tests/test_gpu_traces.py -m gpu -q -s > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|^Class_AMG|passed|failed|error|Error" gpurun_out/pytest_gpu_r2d.log | tail -30
