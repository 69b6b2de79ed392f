# bench.py on N GPUs of one box, as the driver launches it.  Usage: bash tools/gpu_nN_final.sh N
cd $GRAFT_REPO_ROOT
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_final_r2.json 2> gpurun_out/bench_n${N}_final_r2.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_final_r2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['breakdown_ms'], d.get('collectives_per_step'), d['plan_operators'])
r=d['roofline']; print(r['kernel'][:40], r['avg_launch_ms'], r['frac'])
for o in d['roofline_other']: print(o['kernel'][:40], o['avg_launch_ms'], o['frac'])
print(d.get('e2e'))
PY
tail -2 gpurun_out/bench_n${N}_final_r2.err | cut -c1-300
