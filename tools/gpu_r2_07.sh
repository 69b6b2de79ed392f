# round 2, call 7: barrier micro-benchmarks; full GPU suite; smoke; bench (no cpu baseline)
cd $GRAFT_REPO_ROOT
timeout 120 python tools/barrier_bench.py > gpurun_out/barrier_bench_r2g.log 2>&1; echo "barrier bench rc=$?"; cat gpurun_out/barrier_bench_r2g.log
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|passed|failed|rror" gpurun_out/pytest_gpu_r2g.log | tail -20
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_r2g.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2g.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
head -c 900 gpurun_out/bench_r2g.json; echo; tail -3 gpurun_out/bench_r2g.err
