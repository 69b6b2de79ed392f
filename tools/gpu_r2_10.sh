# round 2, call 10: DSMEM cluster solve kernel v2 (register-resident rows): per-op cycles, A/B, tests, bench
cd $GRAFT_REPO_ROOT
SSN_LIB_PATH=$PWD/codes-of-ipd-ssn-amg-method_b200/libssnamg_dbg.so timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 > gpurun_out/amg_prof_dsm_dbg_r2j.log 2>&1; echo "amg_prof dbg rc=$?"
grep -E "pdbg|k30_s1|rror" gpurun_out/amg_prof_dsm_dbg_r2j.log | tail -24
for v in 0 1; do
  SSN_DSM_NOREG=$v timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 5 prof > gpurun_out/amg_prof_noreg${v}_r2j.log 2>&1; echo "amg_prof noreg=$v rc=$?"
  grep -E "k30_s1|solve\.|amg_setup total  |rror" gpurun_out/amg_prof_noreg${v}_r2j.log | tail -9
done
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2j.log 2>&1; echo "pytest rc=$?"
grep -E "^config [0-9]|worst step|solve kernel|passed|failed|rror" gpurun_out/pytest_gpu_r2j.log | tail -30
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc=$?"
head -c 1200 gpurun_out/bench_r2j.json; echo; tail -3 gpurun_out/bench_r2j.err
