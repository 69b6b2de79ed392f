# round 2, call 16 (N GPUs): bench.py under torchrun at N ranks; usage: bash tools/gpu_r2_16.sh N [solve256]
cd $GRAFT_REPO_ROOT
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err; echo "bench n$N rc=$?"
tail -3 gpurun_out/bench_n${N}_r2.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_r2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['breakdown_ms'], d.get('collectives_per_step'), d.get('plan_operators'))
r=d['roofline']; print(r['kernel'][:40], r['avg_launch_ms'], r['frac'], r.get('batch_ms_host_timed'))
for o in d['roofline_other']: print(o['kernel'][:40], o['avg_launch_ms'], o['frac'])
PY
if [ "$2" = "solve256" ]; then
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/run_sharded_solve.py --grid 256 --inner-solver 5 --max-seconds 150 --verbose > gpurun_out/config5_n${N}_r2.json 2> gpurun_out/config5_n${N}_r2.err; echo "config5 n$N rc=$?"
  grep "SsN\|APD" gpurun_out/config5_n${N}_r2.json | cut -c1-200 | tail -14
  grep -v "SsN\|APD" gpurun_out/config5_n${N}_r2.json | cut -c1-1800 | tail -3
  tail -3 gpurun_out/config5_n${N}_r2.err | cut -c1-300
fi
