#!/bin/bash
# round 2, call 47: both NMS bands in one launch: parity, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for d in 5 1; do timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/r2n_bench_depth$d.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2n_bench_depth$d.json')); print('depth $d', round(d['value'],1), round(d['ms_per_step'],3), d.get('gpu_launches_per_step'), {k:v for k,v in d.get('kernel_breakdown_ms_per_step').items() if 'nms' in k})"; done
