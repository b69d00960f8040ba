#!/bin/bash
# round 2, call 11: state with depth 4 / FPS on 2-CTA clusters / aligned compaction: full GPU suite, bench, ncu launch list + full-set captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c11_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r2c11_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c11_bench_b200.json 2> gpurun_out/r2c11_bench_b200.err; echo "b200 rc=$?"; tail -3 gpurun_out/r2c11_bench_b200.err
timeout 400 python bench.py --steps 20 --warmup 3 --min-seconds 5 --no-cpu-baseline > gpurun_out/r2c11_bench_b200_5s.json 2>/dev/null; echo "b200 5s rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c11_bench_b200.json", "gpurun_out/r2c11_bench_b200_5s.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d.get("gpu_launches_per_step"), d.get("parity_in_bench"), d["roofline"]["frac"])
    except Exception as e:
        print("no bench line", f, e)
PY
bash tools/gpu_profile.sh > gpurun_out/r2c11_profile.log 2>&1; tail -5 gpurun_out/r2c11_profile.log
