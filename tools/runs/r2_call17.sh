#!/bin/bash
# round 2, call 17: whole-forward graph replay + one-launch 3-D IoU under the unmodified script (configs[4], 1 GPU)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_gpu.py tests/test_iou3d_roipool_gpu.py tests/test_eval_rcnn_dropin_gpu.py tests/test_checkpoint_roundtrip_gpu.py tests/test_refeval_golden_gpu.py -m gpu -q 2>&1 | tail -5
timeout 900 python tools/run_config5.py --gpus 1 --arms dropin --out /tmp/config5_out > gpurun_out/r2c17_config5_n1.log 2>&1; echo "config5 rc=$?"
grep -E "^dropin" gpurun_out/r2c17_config5_n1.log | cut -c1-330
cp /tmp/config5_out/record.json gpurun_out/r2c17_config5_n1.json 2>/dev/null
PN2_MODEL_GRAPH=0 timeout 900 python tools/run_config5.py --gpus 1 --arms dropin --out /tmp/config5_out_b > gpurun_out/r2c17_config5_n1_eager.log 2>&1
grep -E "^dropin" gpurun_out/r2c17_config5_n1_eager.log | cut -c1-330
timeout 400 python tools/profile_dropin.py --scenes 1920 > gpurun_out/r2c17_profile_dropin.log 2>&1; echo "profile rc=$?"
