#!/bin/bash
# round 2, call 50: configs[4] on one GPU, all three arms, final kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python tools/run_config5.py --gpus 1 --out /tmp/config5_out > gpurun_out/r2p_config5_n1.log 2>&1; echo "config5 rc=$?"
grep -E "^dropin|^fast|^reference" gpurun_out/r2p_config5_n1.log | cut -c1-300
cp /tmp/config5_out/record.json gpurun_out/r2p_config5_n1.json 2>/dev/null
