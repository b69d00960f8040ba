#!/bin/bash
# round 2, call 3: glue kernels, one-launch RCNN input stage, fused RCNN front chain, run-mask compact epilogue
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_gpu.py tests/test_linear_tc_gpu.py tests/test_stock_reference_gpu.py tests/test_mlp_modules_gpu.py tests/test_refnet_golden_gpu.py tests/test_refeval_golden_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c3_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -12 gpurun_out/r2c3_pytest.log
timeout 120 python tools/diag_fps_prefix.py > gpurun_out/r2c3_fps_diag.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r2c3_fps_diag.log | tail -6
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_b200.json 2> gpurun_out/r2c3_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c3_bench_b200.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c3_bench_b200.json"))
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["execution"].get("eager_ms_per_step"))
    print(d["kernel_breakdown_ms_per_step"]); print(d.get("parity_in_bench"))
except Exception as e:
    print("no bench line", e)
PY
timeout 120 python tools/prof_front.py > gpurun_out/r2c3_prof_front.log 2>&1; echo "prof_front rc=$?"; cat gpurun_out/r2c3_prof_front.log
