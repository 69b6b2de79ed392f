# Round check on one B200: GPU parity tests, the bench line, the ncu launch list of one step and
# ncu --set full captures of the plan-wide kernels (summarised into profiles/ by tools/summarize_profiles.py).
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r1.log
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
head -c 1500 gpurun_out/bench_r1.json; echo
SSN_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:plan_trials_screen_kernel -c 3 -f -o gpurun_out/screen_full_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_full_screen.log 2>&1; echo "ncu screen rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:plan_reduce_kernel -c 2 -f -o gpurun_out/k3_full_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_full_k3.log 2>&1; echo "ncu k3 rc=$?"
