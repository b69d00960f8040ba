#!/bin/bash
# round 2, call 27: A2 handed over per K-block in the transposed fused SA kernel: parity, stopwatch, bench at depth 4 / 5 / 6
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_tc_gpu.py tests/test_mlp_modules_gpu.py tests/test_pn2_ops_gpu.py tests/test_refnet_golden_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/prof_sat.py > gpurun_out/r2d_prof_sat_ksplit.log 2>&1; echo "prof rc=$?"
grep -A13 "align 8: \|dense: " gpurun_out/r2d_prof_sat_ksplit.log | grep -v WHAT | head -64
for d in 4 5 6; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/r2d_bench_ksplit_depth$d.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2d_bench_ksplit_depth$d.json')); print('depth $d', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('parity_in_bench',{}).get('matched'))"
done
