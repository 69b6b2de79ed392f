# round 2, call 45: the 181x181-grid solve (1.07e9-entry plan, the largest that fits one B200) with the two-level inner solver
cd $GRAFT_REPO_ROOT
timeout 500 python tools/run_sharded_solve.py --grid 181 --inner-solver 5 --max-seconds 300 > gpurun_out/solve181_s5_r2.json 2> gpurun_out/solve181_s5_r2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/solve181_s5_r2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('inner_solver','outer_its','converged','rel_kkt','objective','warmup_s','loop_s','ssn_steps','line_search_trials','phase_ms','torch_peak_GB_rank0','E_min_median_max','slab_kernels_rank0')})
PY
tail -3 gpurun_out/solve181_s5_r2.err | cut -c1-300
