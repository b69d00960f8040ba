#!/bin/bash
# round 2, call 10: FPS cluster size 2 vs 4 under the pipelined bench (A/B on one box), depth 3 and 4
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 2 0 2; do
  PN2_FPS_CLUSTER=$c timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_c$c.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2c10_bench_c$c.json')); print('cluster $c depth 3:', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
for c in 0 2; do
  PN2_FPS_CLUSTER=$c timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --depth 4 > gpurun_out/r2c10_bench_c${c}_d4.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2c10_bench_c${c}_d4.json')); print('cluster $c depth 4:', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
