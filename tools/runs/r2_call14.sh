#!/bin/bash
# round 2, call 14: native data path under the unmodified script (configs[4], 1 GPU, dropin + fast), N3 GPU test, sanitizers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c14_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/r2c14_pytest.log
timeout 900 python tools/run_config5.py --gpus 1 --arms dropin,fast --out /tmp/config5_out > gpurun_out/r2c14_config5_n1.log 2>&1; echo "config5 rc=$?"
grep -E "^dropin|^fast" gpurun_out/r2c14_config5_n1.log | cut -c1-420
cp /tmp/config5_out/record.json gpurun_out/r2c14_config5_n1.json 2>/dev/null
timeout 600 python tools/run_config5.py --gpus 1 --arms dropin --workers 8 --out /tmp/config5_out_w8 > gpurun_out/r2c14_config5_n1_w8.log 2>&1; echo "config5 w8 rc=$?"
grep -E "^dropin" gpurun_out/r2c14_config5_n1_w8.log | cut -c1-420
cp /tmp/config5_out_w8/record.json gpurun_out/r2c14_config5_n1_w8.json 2>/dev/null
tools/sanitize.sh 300
