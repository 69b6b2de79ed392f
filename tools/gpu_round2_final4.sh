# round 2, last check of the final code: full GPU suite, smoke, both bench configurations with their CPU baselines
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final4_r2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_gpu_final4_r2.log | tail -6
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_final4_r2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final4_r2.log
timeout 1200 python bench.py > gpurun_out/bench_final4_r2.json 2> gpurun_out/bench_final4_r2.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench_final4_r2.json; echo; tail -2 gpurun_out/bench_final4_r2.err
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_final4_r2.json 2> gpurun_out/bench_class2_final4_r2.err; echo "bench class2 rc=$?"
head -c 300 gpurun_out/bench_class2_final4_r2.json; echo
