#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_iou3d.py > gpurun_out/diag_iou3d.log 2>&1; echo "diag exit $?"; cat gpurun_out/diag_iou3d.log | tail -20
timeout 600 python -m pytest tests/test_linear_tc_gpu.py -q -x > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc exit $?"; tail -30 gpurun_out/pytest_tc.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_linear_tc_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
