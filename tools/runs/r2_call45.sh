#!/bin/bash
# round 2, call 45: final evidence of the round (after the ROI-pooling, compaction, interpolation, small-FPS and ball-query changes)
# ncu launch list + --set full summaries (tools/gpu_profile.sh)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench_b200.json 2>gpurun_out/r2m_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200.json')); print('b200', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('gpu_launches_per_step'), d.get('parity_in_bench',{}).get('matched'), d.get('parity_in_bench',{}).get('total')); print(d.get('kernel_breakdown_ms_per_step')); print(d.get('scan_roofline')); print(d.get('cpu_baseline'))"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2m_bench_reference.json 2>gpurun_out/r2m_bench_reference.err
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_reference.json')); print('reference', round(d['value'],1), round(d['e2e']['value'],1))"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 5 > gpurun_out/r2m_bench_b200_5s.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200_5s.json')); print('5s', round(d['value'],1), d['steps'], d['clocks'])"
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/r2m_bench_b200_depth1.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200_depth1.json')); print('depth1', round(d['value'],1), round(d['ms_per_step'],3))"
for d in 4 6; do timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/r2m_bench_b200_depth$d.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200_depth$d.json')); print('depth $d', round(d['value'],1))"; done
timeout 300 python tools/bench_sa_layer.py gpurun_out/r2m_sa_layer_config3.json > gpurun_out/r2m_sa_layer.log 2>&1; tail -12 gpurun_out/r2m_sa_layer.log
timeout 300 python tools/bench_fps_cluster.py > gpurun_out/r2m_bench_fps_variants.log 2>&1; grep "B=8 16384" gpurun_out/r2m_bench_fps_variants.log | cut -c1-150
bash tools/gpu_profile.sh 2>&1 | tail -8
