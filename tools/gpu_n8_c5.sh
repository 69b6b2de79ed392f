# Config 5 (256x256 grids) on 8 GPUs with the two-grid solver and the slab sparse products: verbose log + JSON record.
cd $GRAFT_REPO_ROOT
timeout 520 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/run_sharded_solve.py --grid 256 --inner-solver 5 --max-seconds 240 --verbose > gpurun_out/sharded256_n8_r2b.json 2> gpurun_out/sharded256_n8_r2b.err; echo "sharded256 n8 rc=$?"
grep "SsN\|APD" gpurun_out/sharded256_n8_r2b.json gpurun_out/sharded256_n8_r2b.err | cut -c1-220 | tail -30
grep -v "SsN\|APD" gpurun_out/sharded256_n8_r2b.json | cut -c1-1500 | tail -3
tail -3 gpurun_out/sharded256_n8_r2b.err | cut -c1-300
