#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu49.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu49.log
timeout 600 python bench.py --steps 30 --warmup 4 > gpurun_out/bench49.json 2> gpurun_out/bench49.err; echo "bench exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench49.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["frac"], d["roofline"]["traffic_source"], d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/bench49.err
