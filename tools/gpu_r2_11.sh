# round 2, call 11: building-block micro-benchmarks of the DSMEM cluster kernel; trace tests
cd $GRAFT_REPO_ROOT
timeout 300 python tools/barrier_bench.py > gpurun_out/barrier_bench_r2k.log 2>&1; echo "barrier bench rc=$?"; cat gpurun_out/barrier_bench_r2k.log
timeout 900 python -m pytest tests/test_gpu_traces.py -m gpu -q -s > gpurun_out/pytest_traces_r2k.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|worst step|passed|failed|rror" gpurun_out/pytest_traces_r2k.log | tail -30
