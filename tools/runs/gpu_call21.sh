#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu21.log
timeout 300 python tools/bq_levels.py 2>&1 | tail -14
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench21.json"))
print(d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"])
print(d["kernel_breakdown_ms_per_step"], d["kernel_ms_per_step_sum"])
PY
tail -3 gpurun_out/bench21.err
