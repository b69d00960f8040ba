#!/bin/bash
# round 2, call 32: two-launch compaction + one-launch score argsort: whole suite, bench, ncu launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_b200.json 2>gpurun_out/r2f_bench_b200.err
python -c "import json; d=json.load(open('gpurun_out/r2f_bench_b200.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('gpu_launches'), d.get('parity_in_bench',{}).get('matched')); print(d.get('kernel_breakdown_ms_per_step'))"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/r2f_bench_ncu.log 2>&1; echo "ncu launches exit $?"
grep -c "gpu__time_duration" gpurun_out/r2f_launches.csv
