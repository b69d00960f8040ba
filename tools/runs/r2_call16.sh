#!/bin/bash
# round 2, call 16: shared-memory batch transport under the unmodified script (configs[4], 1 GPU), transposed-small default, racecheck with the longer wait bounds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_tc_gpu.py tests/test_mlp_modules_gpu.py tests/test_refnet_golden_gpu.py tests/test_stock_reference_gpu.py tests/test_eval_rcnn_dropin_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python tools/run_config5.py --gpus 1 --arms dropin --out /tmp/config5_out > gpurun_out/r2c16_config5_n1.log 2>&1; echo "config5 rc=$?"
grep -E "^dropin" gpurun_out/r2c16_config5_n1.log | cut -c1-330
cp /tmp/config5_out/record.json gpurun_out/r2c16_config5_n1.json 2>/dev/null
PN2_SHARED_BATCHES=0 timeout 900 python tools/run_config5.py --gpus 1 --arms dropin --out /tmp/config5_out_b > gpurun_out/r2c16_config5_n1_pickle.log 2>&1
grep -E "^dropin" gpurun_out/r2c16_config5_n1_pickle.log | cut -c1-330
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c16_bench_b200.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2c16_bench_b200.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'])"
tools/sanitize.sh 420
