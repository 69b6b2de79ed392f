set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_plan.py -m gpu -x -q > gpurun_out/pytest_plan_q.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_plan_q.log
for NT in 128 32; do
SSN_LS_MAXNT=$NT timeout 600 python bench.py --no-cpu-baseline --no-full-solve > gpurun_out/bench_q$NT.json 2> gpurun_out/bench_q$NT.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q$NT.json'))
print(d['value'], d['breakdown_ms'], d['config']['line_search_passes'])
print(d['roofline']['kernel'][:40], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline'].get('surviving_entries'), d['roofline'].get('first_pass_NT1_ms'))
for o in d['roofline_other']: print(o['kernel'][:40], o['avg_launch_ms'], o['frac'], o.get('surviving_entries'))
PY
done
timeout 300 python tools/amg_state_prof.py 128 30 > gpurun_out/amg_state_prof.log 2>&1; echo "amgprof rc=$?"
