# round 2, call 24: class 2 phase profile; SpGEMM sort A/B through the AMG profile; AMG tests
cd $GRAFT_REPO_ROOT
timeout 600 python tools/class2_prof.py > gpurun_out/class2_prof_r2w.log 2>&1; grep -v "#launches" gpurun_out/class2_prof_r2w.log | tail -34
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 6 prof > gpurun_out/amg_prof_r2w.log 2>&1; echo "amg_prof rc=$?"
grep -E "k30_s1|amg_setup total  |galerkin|interp W2|rror" gpurun_out/amg_prof_r2w.log | grep -v "#launches" | tail -12
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q > gpurun_out/pytest_amg_r2w.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_amg_r2w.log | tail -6
