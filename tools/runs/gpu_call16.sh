#!/bin/bash
# 2-GPU call: bench at N=2 (torchrun) and the unmodified eval_rcnn.py scene-sharded over 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python - <<'PY' > gpurun_out/eval_sharded.log 2>&1
import importlib, os, subprocess, sys, tempfile, torch
sys.path.insert(0, os.getcwd())
P = "3d_adapt_auto_driving_b200"
et, sk, inf, tu = (importlib.import_module(P + "." + m) for m in ("evaltree", "synthetic_kitti", "inference", "train_utils"))
tmp = tempfile.mkdtemp()
root = et.make_eval_tree(tmp, "oracle/_ref/eval_rcnn.py")
sk.make_dataset(root, n_scenes=10)
model = inf.build_model(seed=0, device="cuda")
with torch.no_grad():
    model.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
os.makedirs(tmp + "/ckpt")
tu.save_checkpoint(tu.checkpoint_state(model, None, 1, 1), filename=tmp + "/ckpt/checkpoint_epoch_1")
common = ["--cfg_file", "cfgs/default.yaml", "--eval_mode", "rcnn", "--ckpt", tmp + "/ckpt/checkpoint_epoch_1.pth", "--batch_size", "1", "--workers", "0"]
# single process reference run (batch 1: per-scene np.random draws do not depend on the shard)
r1 = subprocess.run([sys.executable, "eval_rcnn.py"] + common + ["--output_dir", tmp + "/single"], cwd=root + "/tools", capture_output=True, text=True,
                    env=dict(os.environ, PN2_PER_SCENE_SEED="1"))
print("single rc", r1.returncode, r1.stderr[-500:])
r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29512",
                     "tools/eval_sharded.py", "--tree", root, "--output_dir", tmp + "/sharded", "--"] + common, capture_output=True, text=True)
print("sharded rc", r2.returncode, r2.stdout[-300:], r2.stderr[-800:])
a = tmp + "/single/eval/epoch_1/val/final_result/data"; b = tmp + "/sharded/merged/final_result/data"
fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
print("files", len(fa), len(fb), fa == fb)
same = [open(os.path.join(a, f)).read() == open(os.path.join(b, f)).read() for f in fa]
print("identical files:", sum(same), "/", len(same), " detections:", sum(len(open(os.path.join(a, f)).read().splitlines()) for f in fa))
PY
tail -8 gpurun_out/eval_sharded.log
