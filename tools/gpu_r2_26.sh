# round 2, call 26: pinned ring for small uploads: Hybrid_AMG timing, bench
cd $GRAFT_REPO_ROOT
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 > gpurun_out/amg_prof_r2y.log 2>&1; echo "amg_prof rc=$?"
grep -E "k30_s1|rror" gpurun_out/amg_prof_r2y.log | tail -8
timeout 900 python bench.py --no-cpu-baseline --no-full-solve > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2y.json') if l.startswith('{')][-1])
print(d['value'], d['breakdown_ms'], d['e2e'], d['dominant_by_time'])
PY
tail -3 gpurun_out/bench_r2y.err
