#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_fused" -s 1 -c 1 -f -o gpurun_out/prof_tc3 python tools/prof_tc.py 1 > gpurun_out/prof_tc3.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/prof_tc3.log
