#!/bin/bash
# GPU data path (scene_prepare.cu + gpu_loader.py): parity tests, then end-to-end throughput of tools/eval_fast.py
# (.bin files on disk -> KITTI result files) with the GPU and with the numpy data path on a 480-scene synthetic tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader_gpu.py tests/test_gpu_loader_cpu.py -q -x 2>&1 | tail -15
timeout 900 python - <<'PY' 2>&1 | tail -12
import importlib, os, sys, tempfile, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tools")
PKG = "3d_adapt_auto_driving_b200"
sk = importlib.import_module(PKG + ".synthetic_kitti")
import eval_fast
root = tempfile.mkdtemp()
t0 = time.time()
data_root = sk.make_dataset(root, n_scenes=480, npoints=100000)
print("tree: 480 scenes x 100000 raw points in %.1f s" % (time.time() - t0))
out = {}
for name, kw in (("gpu_loader_warm", dict(gpu_loader=True)), ("gpu_loader", dict(gpu_loader=True)),
                 ("gpu_loader_per_scene_seed", dict(gpu_loader=True, per_scene_seed=True)), ("cpu_loader", dict(gpu_loader=False))):
    os.environ.pop("PN2_PER_SCENE_SEED", None)
    r = eval_fast.run(data_root, os.path.join(root, "out_" + name), batch_size=16, depth=3, **kw)
    out[name] = {k: r[k] for k in ("scenes", "detections", "seconds", "scenes_per_s")}
json.dump(out, open("gpurun_out/eval_fast32.json", "w"), indent=1)
print(json.dumps(out))
PY
