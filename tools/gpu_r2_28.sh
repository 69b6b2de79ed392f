# round 2, call 30: kernel times of the setup after the sort / MIS changes; AMG tests; step timing
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_amg_r2ab.csv python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 2 > gpurun_out/ncu_launches_amg_r2ab.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,re,collections
lines=[l for l in open('gpurun_out/launches_amg_r2ab.csv') if not l.startswith('==')]
tot=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',row['Kernel Name'])[-48:]
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']; v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    tot[name][0]+=1; tot[name][1]+=v
print(sum(v[1] for v in tot.values()), sum(v[0] for v in tot.values()))
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][1])[:12]: print(f"{k:50s} {v[0]:4d} {v[1]:9.1f} {v[1]/v[0]:7.1f}")
PY
timeout 240 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 > gpurun_out/amg_prof_r2ab.log 2>&1; grep -E "k30_s1" gpurun_out/amg_prof_r2ab.log | tail -4
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -q > gpurun_out/pytest_amg_r2ab.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror" gpurun_out/pytest_amg_r2ab.log | tail -4
