# round 2, call 20: DSMEM cluster solve kernel v3 (local fast path, one-barrier reduction): per-op cycles, timing, solver tests
cd $GRAFT_REPO_ROOT
SSN_LIB_PATH=$PWD/codes-of-ipd-ssn-amg-method_b200/libssnamg_dbg.so timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 > gpurun_out/amg_prof_dsm_dbg_r2t.log 2>&1; echo "amg_prof dbg rc=$?"
grep -E "pdbg|rror" gpurun_out/amg_prof_dsm_dbg_r2t.log | tail -24
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_r2t.log 2>&1; echo "amg_prof rc=$?"
grep -E "k30_s1|solve\.|amg_setup total  |rror" gpurun_out/amg_prof_r2t.log | tail -9
timeout 300 python tools/barrier_bench.py 2>&1 | grep -E "z_sum1|z_barrier" 
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -s -k "cluster_resident or three_solve or full_size or class_amg_matches" > gpurun_out/pytest_dsm_r2t.log 2>&1; echo "pytest dsm rc=$?"
grep -E "solve kernel|passed|failed|rror" gpurun_out/pytest_dsm_r2t.log | tail -12
