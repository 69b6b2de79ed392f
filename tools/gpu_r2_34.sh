# round 2, call 34: phase profile of Hybrid_AMG at the benchmarked state; POT bordered solves incl. 'twogrid'
cd $GRAFT_REPO_ROOT
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 6 prof > gpurun_out/amg_prof_phases_r2af.log 2>&1; echo "prof rc=$?"
grep -vE "pdbg|dbg " gpurun_out/amg_prof_phases_r2af.log | head -120
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q -x -k "pot" > gpurun_out/pytest_r2af.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r2af.log
