cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_plan.py -m gpu -x -q 2>&1 | tail -3
python tools/microbench_screen.py 128 16384
python tools/microbench_screen.py 128 2048
timeout 600 python bench.py --no-cpu-baseline --no-full-solve > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print(d['value'], d['breakdown_ms'], d['config']['line_search_passes'])
print(json.dumps(d['roofline'])[:900])
print(d.get('dominant_by_time'))
PY
