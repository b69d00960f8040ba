#!/bin/bash
# round 2, call 35: compute-sanitizer over the kernels added in this session
cd $GRAFT_REPO_ROOT
tools/sanitize_new.sh 200
for t in memcheck synccheck racecheck; do echo "== $t"; sed -n 1,6p gpurun_out/sanitizer_new_$t.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_new_$t.log | head -5; done
