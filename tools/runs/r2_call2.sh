#!/bin/bash
# round 2, call 2: all GPU tests (FPS prefix check, stock reference), sanitizers (racecheck split), ncu source-level captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c2_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/r2c2_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_b200.json 2> gpurun_out/r2c2_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c2_bench_b200.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/r2c2_bench_b200_depth1.json 2>/dev/null; echo "b200 depth1 rc=$?"
tools/sanitize.sh 240
# source-level profile of the two big fused SA launches (grid 592) and the RCNN front linear launches
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sa_fused_t_tc_kernel" --profile-from-start off \
   --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2c2_sa_t python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/ncu1.log 2>&1; echo "ncu sa_t rc=$?"
ncu -i gpurun_out/r2c2_sa_t.ncu-rep --page raw --csv > gpurun_out/r2c2_sa_t_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
cut -c1-300 gpurun_out/r2c2_bench_b200.json
