#!/bin/bash
# transposed-last-layer fused SA kernel: isolated timing + equality vs the row-major fused kernel, then parity + bench
mkdir -p gpurun_out
timeout 300 python tools/prof_tc.py 2 2>&1 | head -4
timeout 900 python -m pytest tests/test_linear_tc_gpu.py tests/test_mlp_modules_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --depth 3 > gpurun_out/bench28.json 2> gpurun_out/bench28.err; echo "bench exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench28.json"))
print(d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"])
print(d["kernel_breakdown_ms_per_step"], d["kernel_ms_per_step_sum"])
PY
tail -3 gpurun_out/bench28.err
