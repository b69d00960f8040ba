#!/bin/bash
# eval_fast sharded over 2 GPUs == single-GPU run with per-scene seeds, byte for byte
mkdir -p gpurun_out
D=$(mktemp -d)
python - <<PY
import importlib, sys
sys.path.insert(0, ".")
sk = importlib.import_module("3d_adapt_auto_driving_b200.synthetic_kitti")
print(sk.make_dataset("$D", n_scenes=96, npoints=40000))
PY
timeout 300 python tools/eval_fast.py --data_root $D/multi_data/kitti --output_dir $D/one --per_scene_seed 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/eval_fast.py --data_root $D/multi_data/kitti --output_dir $D/two 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|^$" | tail -3
python - <<PY
import os
a, b = "$D/one/final_result/data", "$D/two/final_result/data"
fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
same = fa == fb and all(open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read() for f in fa)
nonempty = sum(os.path.getsize(os.path.join(a, f)) > 0 for f in fa)
print("files", len(fa), len(fb), "non-empty", nonempty, "identical:", same)
PY
