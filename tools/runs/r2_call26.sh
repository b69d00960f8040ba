#!/bin/bash
# round 2, call 26: whole GPU suite with the pruned FPS kernel as the default + bench lines (cells vs 2-CTA clusters, same box)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_b200.json 2>gpurun_out/r2d_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2d_bench_b200.json')); print('cells  ', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('parity_in_bench'))"
PN2_FPS_CLUSTER=2 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_b200_fps_cluster2.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2d_bench_b200_fps_cluster2.json')); print('cluster2', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'])"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/r2d_bench_b200_depth1.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2d_bench_b200_depth1.json')); print('depth1 ', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1)); print(d.get('kernel_breakdown_ms_per_step'))"
timeout 300 python tools/bench_fps_cluster.py > gpurun_out/r2d_bench_fps_variants_final.log 2>&1
