#!/bin/bash
# re-entry check of HEAD on a fresh box: GPU parity suite, both bench arms, ncu evidence
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu22.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench22.json 2> gpurun_out/bench22.err; echo "bench exit $?"; cat gpurun_out/bench22.json; tail -3 gpurun_out/bench22.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench22_ref.json 2> gpurun_out/bench22_ref.err; echo "ref exit $?"; cat gpurun_out/bench22_ref.json; tail -3 gpurun_out/bench22_ref.err
bash tools/gpu_profile.sh
