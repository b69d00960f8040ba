#!/bin/bash
# round 2, call 44: unknown points per warp of the culled three_nn (32 / 16 / 8): parity and bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 16 8; do PN2_NN_CPW=$c timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "three_nn or three_interpolate" 2>&1 | tail -1; done
for c in 32 16 8; do
PN2_NN_CPW=$c timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_nn_cpw$c.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2l_bench_nn_cpw$c.json')); print('nn cpw $c', round(d['value'],1), round(d['ms_per_step'],3), {k:v for k,v in d.get('kernel_breakdown_ms_per_step').items() if 'three_nn' in k or 'ball' in k})"
done
