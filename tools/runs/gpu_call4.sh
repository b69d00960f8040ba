#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "max err / scale|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | sort | uniq -c | sort -rn | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench exit $?"
cat gpurun_out/bench_b200.json; tail -5 gpurun_out/bench_b200.err
PN2_MLP=ffma timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b200_ffma.json 2> gpurun_out/bench_b200_ffma.err; echo "bench ffma exit $?"
cat gpurun_out/bench_b200_ffma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --minimal > gpurun_out/bench_ncu.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/bench_ncu.log
