# round 2, call 55: the new two-grid comparison test; the Class 2 bench line with the operator-level phases as a median
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q -x -k "twogrid" -s 2>&1 | grep -E "two-grid in the cluster|passed|failed|rror" | tail -8
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_final5_r2.json 2> gpurun_out/bench_class2_final5_r2.err; echo "bench class2 rc=$?"
head -c 250 gpurun_out/bench_class2_final5_r2.json; echo; grep "amg4pot ms" gpurun_out/bench_class2_final5_r2.err
