# Config 5 (256x256 grids) on 8 GPUs: the row-sharded Class 1 solve, verbose log + JSON record.
cd $GRAFT_REPO_ROOT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/run_sharded_solve.py --grid 256 --max-outer 14 --max-seconds 90 --verbose > gpurun_out/sharded256_n8.json 2> gpurun_out/sharded256_n8.err; echo "sharded256 n8 rc=$?"
grep "SsN\|APD" gpurun_out/sharded256_n8.json | cut -c1-200 | tail -12
grep -v "SsN\|APD" gpurun_out/sharded256_n8.json | cut -c1-1500 | tail -3
tail -3 gpurun_out/sharded256_n8.err | cut -c1-300
