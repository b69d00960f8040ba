#!/bin/bash
# round 2, call 12: BASELINE configs[4] on ONE GPU (7481 scenes, three arms) + the GPU stat_norm option tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stat_norm_gpu.py -m gpu -q 2>&1 | tail -4
nproc; free -g | head -2
timeout 1500 python tools/run_config5.py --gpus 1 --out /tmp/config5_out > gpurun_out/r2c12_config5_n1.log 2>&1; echo "config5 rc=$?"
tail -8 gpurun_out/r2c12_config5_n1.log
cp /tmp/config5_out/record.json gpurun_out/r2c12_config5_n1.json 2>/dev/null
