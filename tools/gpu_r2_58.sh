# round 2, call 58: line search whose first batch reaches past the previous accepted ll (up to 256 steps per read of w):
# the plan / driver / trace tests, then the bench step (no CPU baseline, no full solves)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_driver.py tests/test_gpu_traces.py -m gpu -q -x > gpurun_out/pytest_gpu_r2_58.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2_58.log
timeout 600 python bench.py --no-cpu-baseline --no-full-solve > gpurun_out/bench_r2_58.json 2> gpurun_out/bench_r2_58.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_58.json') if l.startswith('{')][-1])
print(d['value'], d['config'], d['run'], d['breakdown_ms'], d['e2e'])
PY
tail -2 gpurun_out/bench_r2_58.err | cut -c1-300
