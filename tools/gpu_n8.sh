cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,memory.total --format=csv,noheader | head -8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/run_sharded_solve.py --grid 256 --max-outer 14 --max-seconds 90 --verbose > gpurun_out/sharded256_n8.json 2> gpurun_out/sharded256_n8.err; echo "sharded256 n8 rc=$?"
grep -v "^   SsN" gpurun_out/sharded256_n8.json | cut -c1-300 | tail -20
tail -5 gpurun_out/sharded256_n8.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n8b.json') if l.startswith('{')][-1])
print(d['value'], d['breakdown_ms'], d['config']['line_search_passes'], d.get('collectives_per_step'))
print(d['roofline']['kernel'][:40], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline'].get('batch_ms_host_timed'))
PY
