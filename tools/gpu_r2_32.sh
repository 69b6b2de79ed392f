# round 2, call 32: slab SpGEMM on the device (tests + a 131072-node system with a lowered limit), Class 2 tests, class2 bench line
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_driver.py -m gpu -q -x -k "spgemm or class2 or Class2" > gpurun_out/pytest_r2ad.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|rror|skipped" gpurun_out/pytest_r2ad.log | tail -5
timeout 400 python tools/amg_synth.py 256 13 0 > gpurun_out/synth_g256_r13_default.log 2>&1; echo "synth default rc=$?"; grep -E "Hybrid_" gpurun_out/synth_g256_r13_default.log
timeout 400 python tools/amg_synth.py 256 13 0 27 > gpurun_out/synth_g256_r13_slab27.log 2>&1; echo "synth slab rc=$?"; grep -E "Hybrid_" gpurun_out/synth_g256_r13_slab27.log
timeout 900 python bench.py --config class2_64 > gpurun_out/bench_class2_r2ad.json 2> gpurun_out/bench_class2_r2ad.err; echo "bench class2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_class2_r2ad.json')); print(d['value'], d['breakdown_ms'], d['e2e'], d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('same_step_as_device'))"
