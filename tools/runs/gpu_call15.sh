#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_eval_rcnn_dropin_gpu.py tests/test_rotate_iou_gpu.py -q > gpurun_out/pytest_dropin.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_dropin.log | cut -c1-400
