#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_tc_gpu.py -q -x > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc exit $?"; tail -3 gpurun_out/pytest_tc.log
timeout 300 python tools/prof_tc.py 3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench exit $?"
cat gpurun_out/bench_b200.json; tail -5 gpurun_out/bench_b200.err
