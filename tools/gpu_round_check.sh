set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1c.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_r1c.log
timeout 900 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
cat gpurun_out/bench6.json | head -c 3000
SSN_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:plan_trials_lin_kernel -c 3 -f -o gpurun_out/trials_lin_full_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_full_trials_lin.log 2>&1; echo "ncu full rc=$?"
