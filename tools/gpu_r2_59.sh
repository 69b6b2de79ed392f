# round 2, call 59: ncu launch list of one step of the FINAL code (one-read line search)
cd $GRAFT_REPO_ROOT
SSN_BENCH_PROFILE=1 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r2d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_step_r2d.log 2>&1; echo "ncu launch list rc=$?"
wc -l gpurun_out/launches_step_r2d.csv
