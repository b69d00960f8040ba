#!/bin/bash
# round 2, call 38: inverted-rank table + inline exact paths in the pruned FPS kernel: parity, tie-heavy clouds, stopwatch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "fps" 2>&1 | tail -3
timeout 300 python tools/bench_sa_layer.py gpurun_out/r2i_sa_layer_config3.json 2>&1 | grep -E "^lidar|^uniform|^ties" | cut -c1-330
timeout 300 python tools/prof_fps_cells.py lidar 2>&1 | head -8
timeout 300 python tools/prof_fps_cells.py ties 2>&1 | head -8
