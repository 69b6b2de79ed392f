# round 2, call 48: the context's block cache (ssn_ctx::buf_cache) against cudaMallocAsync / cudaFreeAsync pairs: Hybrid_AMG at the
# benchmarked state, then the whole GPU suite with the cache on
cd $GRAFT_REPO_ROOT
for v in 1 0 1 0; do echo "== SSN_BUF_CACHE=$v"; SSN_BUF_CACHE=$v timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 8 2>&1 | grep "k30_s1" | tail -4; done
echo "== profile with the cache"; timeout 300 python tools/amg_prof.py tests/golden/ssn_states_g128.npz k30_s1 4 prof 2>&1 | grep -E "amg_setup total|device buffers|transfer|hybrid\." 
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2_48.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2_48.log
