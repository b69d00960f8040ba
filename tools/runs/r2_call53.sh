#!/bin/bash
# round 2, call 53: the unmodified script in --eval_mode rpn on the drop-in tree
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_eval_rcnn_dropin_gpu.py -m gpu -q -k rpn_mode 2>&1 | tail -25 | cut -c1-250
