# round 2, call 2: cluster-resident solve kernel (parity + time), new parity tests, the bench state fixture
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?"
grep -E "config [0-9]|passed|failed|error" gpurun_out/pytest_gpu_r2b.log | tail -12
SSN_CLUSTER_SOLVE=0 timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof_grid_r2b.log 2>&1; echo "amg prof (grid kernel) rc=$?"
timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof_cluster_r2b.log 2>&1; echo "amg prof (cluster kernel) rc=$?"
SSN_CLUSTER_CTAS=8 timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof_cluster8_r2b.log 2>&1; echo "amg prof (cluster of 8) rc=$?"
for f in grid cluster cluster8; do echo "== $f"; grep "prof=False" gpurun_out/amg_state_prof_${f}_r2b.log | tail -1; grep -E "solve\.(persist|cluster)_solve_kernel  |amg_setup total  |build_dense" gpurun_out/amg_state_prof_${f}_r2b.log; done
timeout 600 python tools/save_bench_state.py 128 30 > gpurun_out/save_bench_state.log 2>&1; echo "save state rc=$?"; tail -2 gpurun_out/save_bench_state.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
head -c 400 gpurun_out/bench_r2b.json; echo
