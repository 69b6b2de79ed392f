# bench.py on N GPUs of one box (N = 1: also the plan-operator GPU tests).  Usage: bash tools/gpu_nN.sh N
cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_plan.py -m gpu -x -q 2>&1 | tail -1
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo "bench n1 rc=$?"
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}c.json 2> gpurun_out/bench_n${N}c.err; echo "bench n$N rc=$?"
fi
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}c.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['breakdown_ms'], d['config']['line_search_passes'], d.get('collectives_per_step'))
r=d['roofline']; print(r['kernel'][:40], r['avg_launch_ms'], r['frac'], r.get('batch_ms_host_timed'))
for o in d['roofline_other']: print(o['kernel'][:40], o['avg_launch_ms'], o['frac'])
PY
