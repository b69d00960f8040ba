#!/bin/bash
# wait loop not unrolled, FAST/PROF as kernel template parameters, MMA warp converged with elect_one
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_tc_gpu.py tests/test_mlp_modules_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/prof_tc.py 2 2>&1 | tail -22
timeout 600 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --depth 3 > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "bench exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench25.json"))
print(d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"])
print(d["kernel_breakdown_ms_per_step"], d["kernel_ms_per_step_sum"])
print(d["roofline"])
PY
tail -3 gpurun_out/bench25.err
