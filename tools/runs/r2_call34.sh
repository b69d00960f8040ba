#!/bin/bash
# round 2, call 34: argsort with -0 == +0
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_glue_gpu.py tests/test_mlp_modules_gpu.py -m gpu -q 2>&1 | grep -E "assert|Error|passed|failed|differ" | head -20
