# round 2, the last call: cycles of the in-kernel PCG phases with the final code (debug build), the 181x181 solve, then the whole check
# of the final code: full GPU suite, smoke, the bench line with its CPU baseline
cd $GRAFT_REPO_ROOT
SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg_dbg.so timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 2>&1 | grep -E "k30_s1|pdbg"
timeout 500 python tools/run_sharded_solve.py --grid 181 --inner-solver 5 --max-seconds 200 > gpurun_out/solve181_s5_r2c.json 2> gpurun_out/solve181_s5_r2c.err; echo "grid 181 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/solve181_s5_r2c.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('inner_solver','outer_its','converged','rel_kkt','objective','warmup_s','loop_s','ssn_steps','phase_ms')})
PY
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final5_r2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_gpu_final5_r2.log | tail -6
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_final5_r2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final5_r2.log
timeout 1200 python bench.py > gpurun_out/bench_final5_r2.json 2> gpurun_out/bench_final5_r2.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench_final5_r2.json; echo; tail -2 gpurun_out/bench_final5_r2.err
