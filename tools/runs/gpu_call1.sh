#!/bin/bash
# first GPU contact: parity tests, goldens from the reference kernels, per-op timings
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/make_goldens.py gpurun_out/golden > gpurun_out/goldens.log 2>&1; echo "goldens exit $?"
timeout 900 python tools/microbench.py gpurun_out/microbench.jsonl > gpurun_out/microbench.log 2>&1; echo "microbench exit $?"
tail -40 gpurun_out/microbench.log
