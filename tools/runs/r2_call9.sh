#!/bin/bash
# round 2, call 9: 8-row aligned compaction + group-wise pooling epilogue (A/B against the run-wise one in the same call)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_tc_gpu.py tests/test_mlp_modules_gpu.py tests/test_refnet_golden_gpu.py -m gpu -q 2>&1 | tail -4
timeout 200 python tools/prof_sat.py > gpurun_out/r2c9_prof_sat.log 2>&1; echo "prof_sat rc=$?"; grep -E "ms \(incl|MMA warp total|EPI E|PROD wait free|wait acc3 free" gpurun_out/r2c9_prof_sat.log
timeout 120 python tools/prof_front.py > gpurun_out/r2c9_prof_front.log 2>&1; echo "prof_front rc=$?"; head -3 gpurun_out/r2c9_prof_front.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench_b200.json 2> gpurun_out/r2c9_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c9_bench_b200.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c9_bench_b200.json"))
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["execution"].get("eager_ms_per_step"))
    print(d["kernel_breakdown_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
