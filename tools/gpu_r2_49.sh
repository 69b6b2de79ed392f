# round 2, call 49: twogrid_bigph's iteration loop inside the cluster kernel (coarse PCG in distributed shared memory) against the
# kernel-by-kernel loop with the grid-wide pcg_kernel; Hybrid_AMG unchanged; the two-grid GPU tests
cd $GRAFT_REPO_ROOT
echo "== kernel by kernel (SSN_TG_CLUSTER=0)"; SSN_TG_CLUSTER=0 SSN_TG_SAVE=/tmp/tg_ref.pt timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 2>&1 | grep -E "k30_s1|twogrid solve|amg_setup total  |pcg|solve\." 
echo "== cluster kernel"; SSN_TG_COMPARE=/tmp/tg_ref.pt timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 2>&1 | grep -E "k30_s1|twogrid solve|amg_setup total  |solve\.|max rel"
echo "== k80_s2"; for v in 0 1; do SSN_TG_CLUSTER=$v timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k80_s2 4 2>&1 | grep -E "k80_s2" | tail -2; done
echo "== Hybrid_AMG (unchanged kernel path)"; timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 6 2>&1 | grep "k30_s1" | tail -3
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_amg.py tests/test_gpu_driver.py tests/test_gpu_traces.py -m gpu -q -x > gpurun_out/pytest_gpu_r2_49.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2_49.log
