#!/bin/bash
# full validation of the round's state: GPU parity suite, both bench arms, ncu evidence
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu30.log
timeout 600 python bench.py --steps 30 --warmup 4 > gpurun_out/bench30.json 2> gpurun_out/bench30.err; echo "bench exit $?"; cat gpurun_out/bench30.json; tail -3 gpurun_out/bench30.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench30_ref.json 2> gpurun_out/bench30_ref.err; echo "ref exit $?"; cat gpurun_out/bench30_ref.json | cut -c1-400; tail -3 gpurun_out/bench30_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_profile.sh
