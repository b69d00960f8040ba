#!/bin/bash
# round 2, call 7: source-level ncu capture of the fused SA kernels (compact mode) after the run-mask epilogue
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sa_fused_t_tc_kernel" --profile-from-start off \
   --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2c7_sa_t python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/ncu7.log 2>&1; echo "ncu sa_t rc=$?"
ncu -i gpurun_out/r2c7_sa_t.ncu-rep --page source --csv --print-source=sass --launch-skip 0 --launch-count 1 > gpurun_out/r2c7_sa_t_sass.csv 2>/dev/null
ncu -i gpurun_out/r2c7_sa_t.ncu-rep --page raw --csv > gpurun_out/r2c7_sa_t_raw.csv 2>/dev/null
rm -f gpurun_out/r2c7_sa_t.ncu-rep
ls -la gpurun_out/r2c7*
