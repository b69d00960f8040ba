#!/bin/bash
# round 2, call 24: pruned FPS kernel v4 (per-cell records)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "fps" -x 2>&1 | tail -4
timeout 300 python tools/prof_fps_cells.py lidar > gpurun_out/r2d_prof_fps_cells_v4.log 2>&1; echo "prof rc=$?"
cat gpurun_out/r2d_prof_fps_cells_v4.log
timeout 300 python tools/bench_fps_cluster.py 2>&1 | grep -E "cells|cluster  4|cluster  1" > gpurun_out/r2d_bench_fps_variants_v4.log; cat gpurun_out/r2d_bench_fps_variants_v2.log
