#!/bin/bash
# fused-SA experiments (epilogue parts switched off), depth 3/4 bench, one ncu full capture of the fused SA kernel with source page
mkdir -p gpurun_out /tmp/ncu
timeout 300 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -x -k "fps" 2>&1 | tail -3
timeout 300 python tools/prof_tc.py 2 2>&1 | tail -22
for d in 3 4; do
timeout 600 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --depth $d > gpurun_out/bench24_d$d.json 2> gpurun_out/bench24_d$d.err; echo "bench depth $d exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench24_d$d.json"))
print(d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"])
print(d["kernel_breakdown_ms_per_step"], d["kernel_ms_per_step_sum"])
PY
tail -3 gpurun_out/bench24_d$d.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sa_fused_tc_kernel" -s 2 -c 1 -f -o /tmp/ncu/fused python tools/prof_tc.py 1 > gpurun_out/ncu_fused.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/ncu/fused.ncu-rep --page raw --csv > gpurun_out/fused_raw.csv 2>/dev/null
ncu -i /tmp/ncu/fused.ncu-rep --page source --csv > gpurun_out/fused_source.csv 2>/dev/null
ls -la gpurun_out/fused_*
