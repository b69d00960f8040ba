#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rotate_iou_gpu.py -q > gpurun_out/pytest_riou.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_riou.log | cut -c1-300
