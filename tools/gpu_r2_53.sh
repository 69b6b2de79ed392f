# round 2, call 53: grid-wide solve kernel with eight loads per lane in flight on the staged dense levels and the masked tails:
# Class 2 bench state (phase profile, then the bench line), the solver / trace tests
cd $GRAFT_REPO_ROOT
timeout 300 python tools/class2_prof.py 2>&1 | grep -E "levels|solve\.persist|amg_setup total  |\{"
timeout 900 python bench.py --config class2_64 --no-cpu-baseline > gpurun_out/bench_class2_r2_53.json 2> gpurun_out/bench_class2_r2_53.err; echo "bench class2 rc=$?"; head -c 250 gpurun_out/bench_class2_r2_53.json; echo
timeout 300 python tools/amg_synth.py 128 9 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_amg.py tests/test_gpu_driver.py tests/test_gpu_traces.py -m gpu -q -x > gpurun_out/pytest_gpu_r2_53.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2_53.log
