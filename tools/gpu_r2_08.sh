# round 2, call 8: polled reads A/B, line-search first batch, traces, bench, ncu --set full of the cluster solve kernel
cd $GRAFT_REPO_ROOT
for v in 0 1; do
  SSN_POLL_READS=$v timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_poll${v}_r2h.log 2>&1; echo "amg_prof poll=$v rc=$?"
  grep -E "k30_s1|cluster_solve_kernel  |amg_setup total  |hybrid" gpurun_out/amg_prof_poll${v}_r2h.log | tail -6
done
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|passed|failed|rror" gpurun_out/pytest_gpu_r2h.log | tail -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
head -c 1200 gpurun_out/bench_r2h.json; echo; tail -3 gpurun_out/bench_r2h.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_solve_kernel -c 1 -f -o gpurun_out/cluster_solve_r2h python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 1 > gpurun_out/ncu_cluster_r2h.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_cluster_r2h.log
