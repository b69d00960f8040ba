#!/bin/bash
# round 2, call 48: 8 KB candidate lists in the culled ball query (twice the resident warps): parity, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "ball_query or group_compaction" 2>&1 | tail -1
for d in 5 1; do timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/r2o_bench_depth$d.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2o_bench_depth$d.json')); print('depth $d', round(d['value'],1), round(d['ms_per_step'],3), {k:v for k,v in d.get('kernel_breakdown_ms_per_step').items() if 'ball' in k})"; done
