# round 2, call 47: live phase profile of Hybrid_AMG at the benchmarked state (where the 3.5 ms of the setup go)
cd $GRAFT_REPO_ROOT
timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 6 prof > gpurun_out/amg_prof_phases_r2_47.log 2>&1; echo "rc=$?"
tail -80 gpurun_out/amg_prof_phases_r2_47.log
