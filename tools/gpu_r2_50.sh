# round 2, call 50: cycles of the phases of the in-kernel PCG (debug build), then the converged 128x128 and 181x181 solves with the
# two-level inner solver now that its loop is one cluster kernel
cd $GRAFT_REPO_ROOT
SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg_dbg.so timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz k30_s1 3 2>&1 | grep -E "k30_s1|pdbg"
for g in 128 181; do
  timeout 500 python tools/run_sharded_solve.py --grid $g --inner-solver 5 --max-seconds 200 > gpurun_out/solve${g}_s5_r2b.json 2> gpurun_out/solve${g}_s5_r2b.err; echo "grid $g rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/solve${g}_s5_r2b.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('inner_solver','outer_its','converged','rel_kkt','objective','warmup_s','loop_s','ssn_steps','line_search_trials','phase_ms')})
PY
done
