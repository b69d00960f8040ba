#!/bin/bash
# round 2, call 19: pruned one-CTA FPS kernel (csrc/fps_cells.cu): parity + latency / SM-time against the cluster kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "fps" -x 2>&1 | tail -8
timeout 300 python tools/bench_fps_cluster.py > gpurun_out/r2d_bench_fps_variants.log 2>&1; echo "bench rc=$?"
cat gpurun_out/r2d_bench_fps_variants.log
