# round 2, call 3: cluster kernel with one pass per slice (lanes-per-row fit), per-(op, level) cycle counters, native driver,
# the full GPU suite, the reference arm at full size, the bench line
cd $GRAFT_REPO_ROOT
D=codes-of-ipd-ssn-amg-method_b200
SSN_LIB_PATH=$PWD/$D/libssnamg_dbg.so timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 prof > gpurun_out/amg_prof_cluster_dbg_r2c.log 2>&1; echo "amg_prof dbg cluster rc=$?"
SSN_CLUSTER_SOLVE=0 SSN_LIB_PATH=$PWD/$D/libssnamg_dbg.so timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 prof > gpurun_out/amg_prof_grid_dbg_r2c.log 2>&1; echo "amg_prof dbg grid rc=$?"
grep -E "k30_s1|pdbg|cluster_solve_kernel  |persist_solve_kernel  " gpurun_out/amg_prof_cluster_dbg_r2c.log | head -40
grep -E "k30_s1|pdbg|persist_solve_kernel  " gpurun_out/amg_prof_grid_dbg_r2c.log | head -30
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|passed|failed|error|Error" gpurun_out/pytest_gpu_r2c.log | tail -14
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2c.json 2> gpurun_out/bench_ref_r2c.err; echo "reference arm rc=$?"
head -c 1500 gpurun_out/bench_ref_r2c.json; echo
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
head -c 700 gpurun_out/bench_r2c.json; echo; tail -3 gpurun_out/bench_r2c.err
