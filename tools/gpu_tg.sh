cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -x -q -k "twogrid" 2>&1 | tail -15
timeout 200 python tools/amg_synth.py 256 4.0 2>&1 | grep "Hybrid_"
timeout 200 python tools/amg_state_prof.py 128 30 2>&1 | grep "prof=False\|Hybrid_twogrid" | tail -5
