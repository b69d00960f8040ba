#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench exit $?"
cat gpurun_out/bench_b200.json; tail -5 gpurun_out/bench_b200.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_b200_eager.json 2> gpurun_out/bench_b200_eager.err; echo "bench eager exit $?"
cut -c1-400 gpurun_out/bench_b200_eager.json; tail -5 gpurun_out/bench_b200_eager.err
