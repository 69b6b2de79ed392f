# round 2, call 51: the PCG leaf with the long rows' tails staged in shared memory against reading them from L2
cd $GRAFT_REPO_ROOT
for st in k30_s1 k80_s2; do
echo "== $st tails from L2 (SSN_PCG_TAIL=0)"; SSN_PCG_TAIL=0 SSN_TG_SAVE=/tmp/tg_$st.pt timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz $st 5 2>&1 | grep -E "$st" | tail -3
echo "== $st tails staged"; SSN_TG_COMPARE=/tmp/tg_$st.pt timeout 300 python tools/twogrid_prof.py tests/golden/ssn_states_g128.npz $st 5 2>&1 | grep -E "$st|max rel|solve.dsm" | tail -5
done
