#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"linear_tc|sa_fused" -s 4 -c 4 -f -o gpurun_out/prof_tc2 python tools/prof_tc.py 2 > gpurun_out/prof_tc2.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/prof_tc2.log
