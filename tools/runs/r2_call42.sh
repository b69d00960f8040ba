#!/bin/bash
# round 2, call 42: copy-style list kernel with hit counts, interpolation weights by three lanes + float4 channels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_b200.json 2>gpurun_out/r2k_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2k_bench_b200.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('gpu_launches_per_step')); print(d.get('kernel_breakdown_ms_per_step'))"
