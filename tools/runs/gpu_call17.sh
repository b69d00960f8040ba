#!/bin/bash
# culled ball query / three_nn, two-source merge GEMM, split roipool: parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pn2_ops_gpu.py tests/test_linear_tc_gpu.py tests/test_iou3d_roipool_gpu.py tests/test_mlp_modules_gpu.py -m gpu -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu17.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?"; cat gpurun_out/bench17.json; tail -5 gpurun_out/bench17.err
