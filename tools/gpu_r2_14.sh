# round 2, call 14: launch list of one Hybrid_AMG call (setup + solve), ncu --set full of the DSMEM cluster solve kernel
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_amg_r2n.csv python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 2 > gpurun_out/ncu_launches_amg_r2n.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_solve_kernel -c 1 -f -o gpurun_out/dsm_solve_r2n python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 1 > gpurun_out/ncu_dsm_r2n.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_dsm_r2n.log
