# round 2, third final check (after the block cache, the slab tiling and the cluster two-grid kernel): full GPU suite, smoke, both bench
# configurations with their CPU baselines, launch list of one step, ncu --set full of the cluster solve kernel in its two-grid form
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final3_r2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_gpu_final3_r2.log | tail -6
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_final3_r2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final3_r2.log
timeout 1200 python bench.py > gpurun_out/bench_final3_r2.json 2> gpurun_out/bench_final3_r2.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_final3_r2.json; echo; tail -2 gpurun_out/bench_final3_r2.err
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_final3_r2.json 2> gpurun_out/bench_class2_final3_r2.err; echo "bench class2 rc=$?"
head -c 300 gpurun_out/bench_class2_final3_r2.json; echo
SSN_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r2c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_step_r2c.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_solve_kernel -c 1 -s 2 -f -o gpurun_out/dsm_twogrid_full_r2 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k30_s1 4 > gpurun_out/ncu_full_dsm_twogrid_r2.log 2>&1; echo "ncu dsm twogrid rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
