"""Statistical Normalization point rescale: scenes/s of the GPU batch path (stat_norm/gpu_rescale.py) next to the
numpy path (stat_norm/norm.py rescale_ptc + the float32 cast of format_lidar_data) on the same synthetic scenes
(120 000 points, 8 Car / Van boxes each).  python tools/bench_stat_norm.py [out.json]"""
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "3d_adapt_auto_driving_b200"


def main(out_path):
    norm = importlib.import_module(PKG + ".stat_norm.norm")
    gr = importlib.import_module(PKG + ".stat_norm.gpu_rescale")
    o3 = importlib.import_module(PKG + ".stat_norm.object_3d")
    ku = importlib.import_module(PKG + ".stat_norm.kitti_util")
    sk = importlib.import_module(PKG + ".synthetic_kitti")
    d = tempfile.mkdtemp()
    cpath = os.path.join(d, "calib.txt")
    open(cpath, "w").write("\n".join(sk.CALIB_LINES) + "\n")
    calib = ku.Calibration(cpath)
    mapping = norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)
    rng = np.random.RandomState(0)
    import test_stat_norm_gpu as tg
    scenes = []
    for _ in range(32):
        labels = tg._objects(o3, rng, 8)
        scenes.append((tg._scene(calib, labels, rng, 120000), labels, calib))
    dev = torch.device("cuda:0")
    gr.rescale_scenes_gpu(mapping, scenes[:16], dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(5):
        for s in range(0, 32, 16):
            gr.rescale_scenes_gpu(mapping, scenes[s:s + 16], dev)
    torch.cuda.synchronize()
    gpu = 5 * 32 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for velo, labels, c in scenes[:16]:
        pts, _ = norm.rescale_ptc(mapping, velo, labels, c)
        np.concatenate([pts, np.ones((pts.shape[0], 1), dtype=np.float32)], axis=1).astype(np.float32)
    cpu = 16 / (time.perf_counter() - t0)
    res = {"scenes_per_s_gpu_batch16_incl_h2d_d2h": round(gpu, 1), "scenes_per_s_numpy_1_thread": round(cpu, 1),
           "points_per_scene": 120000, "boxes_per_scene": 8}
    print(json.dumps(res))
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stat_norm_bench.json")
