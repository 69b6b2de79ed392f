# round 2, call 46: slab-sized launches of the plan-wide kernels (what a rank of an 8-GPU run sees) against the number of waves
cd $GRAFT_REPO_ROOT
for wv in 2 1 3 4; do for st in 1 0; do SSN_PLAN_WAVES=$wv SSN_PLAN_STAGE=$st timeout 120 python tools/microbench_slab.py 2048 16384 2>&1 | tail -3; done; done
echo "== 8192-row slab (2 GPUs)"; for wv in 2 4; do SSN_PLAN_WAVES=$wv timeout 120 python tools/microbench_slab.py 8192 16384 2>&1 | tail -3; done
