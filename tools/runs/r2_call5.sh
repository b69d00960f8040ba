#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_gpu.py -m gpu -q -s 2>&1 | grep -E "rot_mode|passed|failed|FAILED|differ" | head -20
timeout 120 python tools/prof_front.py > gpurun_out/r2c5_prof_front.log 2>&1; echo "prof_front rc=$?"; cat gpurun_out/r2c5_prof_front.log
