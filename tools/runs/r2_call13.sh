#!/bin/bash
# round 2, call 13 (8 GPUs): BASELINE configs[4] on 8 x B200 (three arms; script-default 4 DataLoader workers per rank, then 16) and the 8-GPU bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | sed -n 2p
timeout 900 python tools/run_config5.py --gpus 8 --out /tmp/config5_out > gpurun_out/r2c13_config5_n8.log 2>&1; echo "config5 n8 rc=$?"
tail -6 gpurun_out/r2c13_config5_n8.log | cut -c1-900
cp /tmp/config5_out/record.json gpurun_out/r2c13_config5_n8.json 2>/dev/null
timeout 600 python tools/run_config5.py --gpus 8 --arms dropin --workers 16 --out /tmp/config5_out_w16 > gpurun_out/r2c13_config5_n8_w16.log 2>&1; echo "config5 n8 w16 rc=$?"
cp /tmp/config5_out_w16/record.json gpurun_out/r2c13_config5_n8_w16.json 2>/dev/null
tail -3 gpurun_out/r2c13_config5_n8_w16.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2c13_bench_b200_n8.json 2> gpurun_out/r2c13_bench_b200_n8.err; echo "bench n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c13_bench_reference_n8.json 2> gpurun_out/r2c13_bench_reference_n8.err; echo "bench ref n8 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c13_bench_b200_n8.json", "gpurun_out/r2c13_bench_reference_n8.json"):
    try:
        line = [l for l in open(f) if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1))
    except Exception as e:
        print("no bench line", f, e)
PY
