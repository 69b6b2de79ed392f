# round 2, call 18: Class 2 fused residual: tests, config-3 bench line; full suite
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_traces.py -m gpu -q -s -k "pot or config3" > gpurun_out/pytest_pot_r2r.log 2>&1; echo "pytest pot rc=$?"
grep -E "^config|passed|failed|rror" gpurun_out/pytest_pot_r2r.log | tail -12
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_r2r.json 2> gpurun_out/bench_class2_r2r.err; echo "bench class2 rc=$?"
head -c 2500 gpurun_out/bench_class2_r2r.json; echo; tail -3 gpurun_out/bench_class2_r2r.err
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2r.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_gpu_r2r.log | tail -10
