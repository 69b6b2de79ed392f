# round 2, call 52: where the grid-wide solve kernel spends a W-cycle on the Class 2 bench state (8192 / 4096 / 21 levels, 1.6M / 3.5M nonzeros)
cd $GRAFT_REPO_ROOT
SSN_LIB_PATH=$GRAFT_REPO_ROOT/codes-of-ipd-ssn-amg-method_b200/libssnamg_dbg.so timeout 300 python tools/class2_prof.py 2>&1 | grep -E "pdbg|levels|solve\.|amg_setup total  |\{" 
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_solve_kernel -c 1 -s 3 -f -o gpurun_out/persist_class2_full_r2 python tools/class2_prof.py > gpurun_out/ncu_full_persist_class2_r2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/persist_class2_full_r2.ncu-rep
