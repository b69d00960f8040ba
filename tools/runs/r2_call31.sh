#!/bin/bash
# round 2, call 31: which association does torch.sum use for three terms?  + the rest of the suite after the failing test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pn2_ops_gpu.py -m gpu -q -k "three_interpolate_from or fps_writes or fill_variant" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
