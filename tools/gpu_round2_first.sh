# First gpurun call of round 2 (one B200): the regular GPU suite, then the device-setup paths that were written after
# round 1's GPU budget was spent (tests/test_zz_device_setup.py, the standalone C program of tests/test_matio.py runs
# with the regular suite), then the bench line.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r2.log
SSN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_zz_device_setup.py -m gpu -q > gpurun_out/pytest_device_setup_r2.log 2>&1; echo "device-setup rc=$?"
tail -15 gpurun_out/pytest_device_setup_r2.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2.log
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_r2.json; echo
# what the end-of-round-1 setup changes bought (measured nowhere yet): Hybrid_AMG at the benchmarked state with the old
# and the new scan threshold (round 1: 12.8 ms, 371 launches; setup 4.6 ms)
SSN_SMALL_SCAN_MAX=262144 timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof_scan262144_r2.log 2>&1; echo "amg prof (old scans) rc=$?"
timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof_r2.log 2>&1; echo "amg prof rc=$?"
grep "prof=False" gpurun_out/amg_state_prof_scan262144_r2.log | tail -1; grep "prof=False" gpurun_out/amg_state_prof_r2.log | tail -1
grep "amg_setup total" gpurun_out/amg_state_prof_scan262144_r2.log gpurun_out/amg_state_prof_r2.log
