#!/bin/bash
# round 2, call 1: GPU tests incl. the stock-reference tests, both bench arms, sanitizer subset
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c1_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/r2c1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c1_smoke.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c1_bench_reference.json 2> gpurun_out/r2c1_bench_reference.err; echo "ref rc=$?"; tail -3 gpurun_out/r2c1_bench_reference.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c1_bench_b200.json 2> gpurun_out/r2c1_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c1_bench_b200.err
python bench.py --steps 20 --warmup 3 --min-seconds 5 --no-cpu-baseline > gpurun_out/r2c1_bench_b200_5s.json 2>/dev/null; echo "b200 5s rc=$?"
tools/sanitize.sh 300
cut -c1-600 gpurun_out/r2c1_bench_reference.json; echo; cut -c1-400 gpurun_out/r2c1_bench_b200.json
