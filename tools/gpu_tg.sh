cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for S in 4 5; do
timeout 200 python tools/run_sharded_solve.py --grid 181 --max-outer 4 --max-seconds 30 --inner-solver $S --verbose > gpurun_out/solve181_s$S.json 2> gpurun_out/solve181_s$S.err; echo "solve181 inner_solver=$S rc=$?"
grep "SsN\|APD" gpurun_out/solve181_s$S.json | cut -c1-150 | tail -14
done
