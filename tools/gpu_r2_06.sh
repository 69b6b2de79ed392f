# round 2, call 6: where the fused setup kernel spends its cycles; class 2 warm start device vs oracle
cd $GRAFT_REPO_ROOT
SSN_FUSED_PROF=1 timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 2 > gpurun_out/fused_prof_r2f.log 2>&1; echo "fused prof rc=$?"
grep -E "fused setup|k30_s1" gpurun_out/fused_prof_r2f.log | head
timeout 900 python tools/debug_class2_warmup.py > gpurun_out/class2_warmup_r2f.log 2>&1; echo "class2 warmup rc=$?"; cat gpurun_out/class2_warmup_r2f.log | tail -20
