# round 2, call 57: compute-sanitizer (memcheck) over the paths changed late in the round: the two-level solver inside the cluster kernel
# (PCG leaf, staged tails), the grid-wide kernel's masked tails (forced with set_cluster_solve(0) in the three-kernel tests), the block cache
cd $GRAFT_REPO_ROOT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_solvers.py -m gpu -q -x -k "twogrid_matches_oracle or twogrid_bigph_matches or generic_twogrid or golden" > gpurun_out/sanitizer_twogrid_r2.log 2>&1; echo "memcheck two-grid rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_twogrid_r2.log | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_amg.py -m gpu -q -x -k "cluster_resident_solve_paths_agree or dense_tail or persistent" > gpurun_out/sanitizer_amg_r2.log 2>&1; echo "memcheck amg rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_amg_r2.log | tail -6
