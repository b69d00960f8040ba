#!/bin/bash
# round 2, call 18: configs[4] on ONE GPU, all three arms, final data path
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_iou3d_roipool_gpu.py tests/test_eval_rcnn_dropin_gpu.py tests/test_gpu_loader_gpu.py -m gpu -q 2>&1 | tail -3
timeout 1500 python tools/run_config5.py --gpus 1 --out /tmp/config5_out > gpurun_out/r2c18_config5_n1.log 2>&1; echo "config5 rc=$?"
grep -E "^dropin|^fast|^reference" gpurun_out/r2c18_config5_n1.log | cut -c1-300
cp /tmp/config5_out/record.json gpurun_out/r2c18_config5_n1.json 2>/dev/null
