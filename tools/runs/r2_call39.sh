#!/bin/bash
# round 2, call 39: three-phase ROI pooling kernel + merged RPN head layers: parity, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_gpu.py tests/test_iou3d_roipool_gpu.py tests/test_mlp_modules_gpu.py tests/test_refnet_golden_gpu.py tests/test_stock_reference_gpu.py -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_b200.json 2>gpurun_out/r2i_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2i_bench_b200.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('gpu_launches_per_step')); print({k:v for k,v in d.get('kernel_breakdown_ms_per_step').items() if 'roipool' in k or 'linear_tc' in k})"
