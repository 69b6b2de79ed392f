# round 2, call 36: cp.async staged reduction kernels (SSN_PLAN_STAGE) A/B, Aty with 8 waves; plan tests
cd $GRAFT_REPO_ROOT
for st in 1 0 1 0; do echo "== SSN_PLAN_STAGE=$st"; SSN_PLAN_STAGE=$st timeout 300 python tools/microbench.py 128 2>&1 | grep -E "^Ax|^Aty|^prox_residual|torch copy"; done
timeout 900 python -m pytest tests/test_gpu_plan.py -m gpu -q -x > gpurun_out/pytest_plan_r2ag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_plan_r2ag.log
