# round 2, call 15: scan with fused publish, one-wave dense build; full GPU suite, smoke, bench with the reference-side fields
cd $GRAFT_REPO_ROOT
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_r2o.log 2>&1; echo "amg_prof rc=$?"
grep -E "k30_s1|solve\.|amg_setup total|rror" gpurun_out/amg_prof_r2o.log | tail -12
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2o.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_gpu_r2o.log | tail -10
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_r2o.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2o.log
timeout 900 python bench.py > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "bench rc=$?"
head -c 1500 gpurun_out/bench_r2o.json; echo; tail -3 gpurun_out/bench_r2o.err
