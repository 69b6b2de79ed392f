# round 2, call 9: the DSMEM cluster solve kernel: profile A/B, new tests, full suite, bench
cd $GRAFT_REPO_ROOT
for v in 1 0; do
  SSN_DSM_SOLVE=$v timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_dsm${v}_r2i.log 2>&1; echo "amg_prof dsm=$v rc=$?"
  grep -E "k30_s1|solve\.|amg_setup total  |rror" gpurun_out/amg_prof_dsm${v}_r2i.log | tail -12
done
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q -s -k "cluster_resident or three_solve or full_size or class_amg_matches" > gpurun_out/pytest_dsm_r2i.log 2>&1; echo "pytest dsm rc=$?"
grep -E "solve kernel|cluster solve mode|Class_AMG|passed|failed|rror" gpurun_out/pytest_dsm_r2i.log | tail -40
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|passed|failed|rror" gpurun_out/pytest_gpu_r2i.log | tail -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"
head -c 1500 gpurun_out/bench_r2i.json; echo; tail -3 gpurun_out/bench_r2i.err
