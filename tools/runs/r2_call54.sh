#!/bin/bash
# round 2, call 54: bench.py after the pruned_fps reporting change (short run)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_check.json 2>gpurun_out/r2q_bench_check.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2q_bench_check.json')); print(round(d['value'],1), d['scan_roofline'], d['pruned_fps'], d['roofline']['frac'], d['gpu_launches'])"
tail -2 gpurun_out/r2q_bench_check.err
