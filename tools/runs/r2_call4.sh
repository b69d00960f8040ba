#!/bin/bash
# round 2, call 4: rot-mode pin of the canonical transform, rcnn_front v2 (direct H stores, staged row heads, M3 before M2F)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_gpu.py -m gpu -q -s 2>&1 | grep -E "rot_mode|passed|failed|FAILED|differ" | head -20
timeout 300 python -m pytest "tests/test_linear_tc_gpu.py::test_rcnn_front_chain_in_one_launch" tests/test_mlp_modules_gpu.py tests/test_refnet_golden_gpu.py -m gpu -q 2>&1 | tail -5
timeout 120 python tools/prof_front.py > gpurun_out/r2c4_prof_front.log 2>&1; echo "prof_front rc=$?"; cat gpurun_out/r2c4_prof_front.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_b200.json 2> gpurun_out/r2c4_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c4_bench_b200.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c4_bench_b200.json"))
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["execution"].get("eager_ms_per_step"))
    print(d["kernel_breakdown_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
