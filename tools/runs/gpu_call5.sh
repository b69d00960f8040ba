#!/bin/bash
mkdir -p gpurun_out
python tools/prof_tc.py 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 2 -f -o gpurun_out/prof_linear_tc python tools/prof_tc.py 2 > gpurun_out/prof_tc.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/prof_tc.log
ls -la gpurun_out/*.ncu-rep
