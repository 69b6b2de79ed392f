# As gpu_round_check.sh without the ncu --set full captures.
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_r1.log
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
SSN_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1.json'))
print(d['value'], d['breakdown_ms'], d['e2e']['value'], d['dominant_by_time'], d['full_solve']['total_s'], d['roofline']['traffic'], d['roofline']['frac'])
PY
