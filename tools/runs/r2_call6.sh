#!/bin/bash
# round 2, call 6: full GPU suite + bench (both arms) on the state with glue kernels, one-launch RCNN input stage, fused front chain
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c6_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/r2c6_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c6_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c6_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c6_bench_b200.json 2> gpurun_out/r2c6_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c6_bench_b200.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c6_bench_reference.json 2> gpurun_out/r2c6_bench_reference.err; echo "ref rc=$?"; tail -3 gpurun_out/r2c6_bench_reference.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c6_bench_b200.json", "gpurun_out/r2c6_bench_reference.json"):
    try:
        d = json.load(open(f))
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"), d.get("parity_in_bench"))
    except Exception as e:
        print("no bench line", f, e)
PY
