# round 2, second final check: full GPU suite, smoke, both bench configurations (with their CPU baselines), launch list of one step,
# ncu --set full captures of the fused residual and of the screen kernel (summarised by tools/summarize_profiles.py r2 / step_r2)
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final2_r2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_gpu_final2_r2.log | tail -6
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_final2_r2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final2_r2.log
timeout 1200 python bench.py > gpurun_out/bench_final2_r2.json 2> gpurun_out/bench_final2_r2.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_final2_r2.json; echo; tail -2 gpurun_out/bench_final2_r2.err
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_final2_r2.json 2> gpurun_out/bench_class2_final2_r2.err; echo "bench class2 rc=$?"
head -c 300 gpurun_out/bench_class2_final2_r2.json; echo
SSN_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_step_r2.log 2>&1; echo "ncu launch list rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:plan_reduce_kernel -c 2 -f -o gpurun_out/k3_full_r2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_full_k3_r2.log 2>&1; echo "ncu k3 rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:plan_trials_screen_kernel -c 2 -f -o gpurun_out/screen_full_r2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_full_screen_r2.log 2>&1; echo "ncu screen rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
