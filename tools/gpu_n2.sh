set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_driver.py -m gpu -x -q -k "sharded" > gpurun_out/pytest_drv_q.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_drv_q.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/run_sharded_solve.py --grid 128 --max-outer 25 > gpurun_out/sharded128_n2.json 2> gpurun_out/sharded128_n2.err; echo "sharded128 n2 rc=$?"
tail -5 gpurun_out/sharded128_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/sharded128_n2.json') if l.startswith('{')][-1])
for k in ("warmup_s","loop_s","outer_its","converged","rel_kkt","objective","ssn_steps","line_search_trials","line_search_passes","phase_ms","collectives","torch_peak_GB_rank0","E_min_median_max"): print(k, d[k])
print(d["fxk"][:8])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2b.json') if l.startswith('{')][-1])
print(d['value'], d['breakdown_ms'], d['config']['line_search_passes'], d.get('collectives_per_step'))
print(d['roofline']['kernel'][:40], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline'].get('batch_ms_host_timed'))
PY
