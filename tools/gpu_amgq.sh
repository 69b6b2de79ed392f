cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_solvers.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/amg_state_prof.py 128 30 2>&1 | grep "prof=False\|persist_solve_kernel  \|amg_setup total  \|solve loop total  \|launches" | tail -8
