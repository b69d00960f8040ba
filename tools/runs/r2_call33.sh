#!/bin/bash
# round 2, call 33: register-blocked argsort: test diagnostics, suite, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_gpu.py -m gpu -q -k "argsort" 2>&1 | grep -E "assert|Error|passed|failed|differ" | head -20
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_b200.json 2>gpurun_out/r2g_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2g_bench_b200.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('gpu_launches')); print({k:v for k,v in d.get('kernel_breakdown_ms_per_step').items() if 'argsort' in k or 'compact' in k})"
