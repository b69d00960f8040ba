#!/bin/bash
# round 2, call 40: two-warp FPS for many small clouds; culled vs brute-force ball query on ROI clouds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "fps" 2>&1 | tail -3
timeout 300 python tools/bench_small_ball_query.py 2>&1 | tail -8
