set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -x -q -k "sharded or full_solve" > gpurun_out/pytest_drv_q.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_drv_q.log
timeout 300 python tools/run_sharded_solve.py --grid 64 > gpurun_out/sharded64_n1.json 2> gpurun_out/sharded64_n1.err; echo "sharded64 rc=$?"
tail -3 gpurun_out/sharded64_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sharded64_n1.json'))
for k in ("warmup_s","loop_s","outer_its","converged","rel_kkt","objective","ssn_steps","line_search_trials","line_search_passes","phase_ms","collectives","torch_peak_GB_rank0"): print(k, d[k])
PY
