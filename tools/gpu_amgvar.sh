cd $GRAFT_REPO_ROOT
for v in default 512_1 1024_1 256_1; do
  if [ $v = default ]; then unset SSN_LIB_PATH; else export SSN_LIB_PATH=$GRAFT_REPO_ROOT/devlibs/libssnamg_$v.so; fi
  echo "== $v"; timeout 200 python tools/amg_state_prof.py 128 30 2>&1 | grep "prof=False\|persist_solve_kernel  \|amg_setup total  \|solve loop total  " | tail -5
done
