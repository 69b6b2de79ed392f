# round 2, call 44: the 128x128 solve with the two-level inner solver (inner_solver = 5) next to the default (4), one GPU
cd $GRAFT_REPO_ROOT
for s in 5 4; do
  timeout 400 python tools/run_sharded_solve.py --grid 128 --inner-solver $s --max-seconds 150 > gpurun_out/solve128_s${s}_r2.json 2> gpurun_out/solve128_s${s}_r2.err; echo "solver $s rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/solve128_s${s}_r2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('inner_solver','outer_its','converged','rel_kkt','objective','warmup_s','loop_s','ssn_steps','line_search_trials','phase_ms')})
PY
done
