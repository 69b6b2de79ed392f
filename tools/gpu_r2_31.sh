# round 2, call 31: Class 2 as library calls (ssn_warmup_class2 / ssn_apd_ssn_class2 / ssn_ssn_step_class2): GPU suite, class2 bench
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2ac.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_gpu_r2ac.log | tail -8
timeout 600 python bench.py --config class2_64 --no-cpu-baseline > gpurun_out/bench_class2_r2ac.json 2> gpurun_out/bench_class2_r2ac.err; echo "bench class2 rc=$?"
head -c 1500 gpurun_out/bench_class2_r2ac.json; echo; tail -3 gpurun_out/bench_class2_r2ac.err
timeout 600 python tools/run_class2.py 64 120 4 > gpurun_out/run_class2_r2ac.log 2>&1; echo "run_class2 rc=$?"; tail -3 gpurun_out/run_class2_r2ac.log
