# round 2, call 21: one-call step: test, bench
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q > gpurun_out/pytest_drv_r2u.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_drv_r2u.log
timeout 900 python bench.py --no-cpu-baseline --no-full-solve > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2u.json') if l.startswith('{')][-1])
print(d['value'], d['breakdown_ms'], d['e2e'], d.get('plan_operators'))
PY
tail -3 gpurun_out/bench_r2u.err
