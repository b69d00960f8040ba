#!/bin/bash
# round 2, call 51: 8-GPU bench line with the final kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_b200_n8.json 2>gpurun_out/r2m_bench_b200_n8.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200_n8.json')); print('n8', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['n_gpus'])"
