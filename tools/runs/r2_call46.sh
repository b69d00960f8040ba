#!/bin/bash
# round 2, call 46: 2-GPU bench line with the final kernels (scene shards + one all_gather)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_b200_n2.json 2>gpurun_out/r2m_bench_b200_n2.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2m_bench_b200_n2.json')); print('n2', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['n_gpus'])"
tail -3 gpurun_out/r2m_bench_b200_n2.err
