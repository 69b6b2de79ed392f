# round 2, call 54: phase profile of the inner solvers on a synthetic early-phase system of the 128x128 solve (E ~ 8.7e6)
cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import sys, runpy
sys.argv = ["amg_synth.py", "128", "13"]
import ssnamg
_dump = ssnamg.profile_dump
def dump_all():
    t = _dump(); print(t); return ""
ssnamg.profile_dump = dump_all
runpy.run_path("tools/amg_synth.py", run_name="__main__")
PY
