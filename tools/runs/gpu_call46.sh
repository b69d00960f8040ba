#!/bin/bash
# final state of the round: full GPU parity suite, both bench arms, ncu evidence (tag r1d)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu46.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu46.log
timeout 600 python bench.py --steps 30 --warmup 4 > gpurun_out/bench46.json 2> gpurun_out/bench46.err; echo "bench exit $?"; cat gpurun_out/bench46.json | cut -c1-300; tail -3 gpurun_out/bench46.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench46_ref.json 2> gpurun_out/bench46_ref.err; echo "ref exit $?"; cat gpurun_out/bench46_ref.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_profile.sh
