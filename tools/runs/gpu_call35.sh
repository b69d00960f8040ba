#!/bin/bash
# evaluator on the GPU path + the complete GPU suite + end-to-end: eval_fast on a synthetic tree -> kitti_ap
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu35.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu35.log
timeout 600 python - <<'PY' 2>&1 | tail -25
import importlib, os, sys, tempfile, time, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tools")
PKG = "3d_adapt_auto_driving_b200"
sk = importlib.import_module(PKG + ".synthetic_kitti")
import eval_fast, torch
inf = importlib.import_module(PKG + ".inference")
orig = inf.build_model
def biased(seed=0, eval_mode="rcnn", device="cuda"):
    m = orig(seed, eval_mode, device)
    with torch.no_grad():
        m.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
    return m
inf.build_model = biased
root = tempfile.mkdtemp()
data_root = sk.make_dataset(root, n_scenes=200, npoints=60000)
r = eval_fast.run(data_root, os.path.join(root, "out"), batch_size=16, depth=3)
t0 = time.time()
p = subprocess.run([sys.executable, "tools/kitti_ap.py", "--label_dir", os.path.join(data_root, "KITTI/object/training/label_2"),
                    "--result_dir", r["final_dir"], "--split_file", os.path.join(data_root, "KITTI/ImageSets/val.txt")],
                   capture_output=True, text=True)
print(p.stdout[-1500:]); print(p.stderr[-600:]); print("kitti_ap wall %.1f s" % (time.time() - t0))
PY
