"""Per-scene HOST costs of the disk -> result-file pipeline (tools/eval_fast.py), measured without a GPU: file reads,
calibration / image-header parsing, the native MT19937 replay of the reference's sampling draws, result formatting.
The GPU forward pass takes ~0.52 ms per scene (bench.py), so these are what bounds eval_fast end to end.

    python tools/bench_host_path.py [--json profiles/<tag>_host_path_cpu.json]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def load(sub):
    return importlib.import_module(PKG + "." + sub)


def per_call_ms(fn, n, warmup=1, repeats=3):
    """best of `repeats` timed loops of n calls (the build container's cores are shared: single loops vary 2x)"""
    for i in range(warmup):
        fn(i)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        best = min(best, (time.perf_counter() - t0) / n * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--raw_points", type=int, default=100000)
    args = ap.parse_args()
    load("config").use_default_yaml("rcnn")
    sk, ko, gl = load("synthetic_kitti"), load("kitti_output"), load("datasets.gpu_loader")
    with tempfile.TemporaryDirectory() as d:
        root = sk.make_dataset(d, name="kitti", n_scenes=16, split="val", seed=1, npoints=args.raw_points)
        ds = load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=root, npoints=16384, split="val", mode="EVAL",
                                                                  classes="Car", far_points=4000)
        res = {"raw_points_per_scene": args.raw_points, "unit": "ms per scene, one host thread"}
        res["read_velodyne_bin"] = per_call_ms(lambda i: ds.get_lidar(i % 16), 64)
        res["parse_calib"] = per_call_ms(lambda i: ds.get_calib(i % 16), 64)
        res["read_image_shape"] = per_call_ms(lambda i: ds.get_image_shape(i % 16), 64)
        calib = ds.get_calib(0)
        rng = np.random.RandomState(0)
        boxes = np.concatenate([rng.uniform(-30, 30, (50, 1)), rng.uniform(-1, 3, (50, 1)), rng.uniform(5, 70, (50, 1)),
                                rng.uniform(1, 2, (50, 3)), rng.uniform(-4, 4, (50, 1))], 1).astype(np.float32)
        scores = rng.randn(50).astype(np.float32)
        res["write_result_file_50_boxes"] = per_call_ms(lambda i: ko.save_kitti_format(i % 16, calib, boxes, d, scores, (375, 1242, 3)), 200)
        # draws of a scene with 50 k valid points, 30 % of them beyond 40 m (the far list is subsampled to 4000)
        n_valid, n_far = 50000, 15000
        st = gl.MTState.seeded(1)
        out = np.empty(16384, np.int32)
        scratch = np.empty(n_valid + 16384, np.int32)
        res["mt19937_draws_native"] = per_call_ms(
            lambda i: gl.draw_selection_native(st, n_valid, n_valid - n_far, n_far, 16384, 4000, False, out, scratch), 200)
        rs = np.random.RandomState(1)
        res["mt19937_draws_numpy"] = per_call_ms(
            lambda i: gl.draw_selection(n_valid, n_valid - n_far, n_far, 16384, 4000, False, rng=rs), 30)
        np.random.seed(0)
        # the first passes are 3-5x slower (allocator growth for the ~20 MB of temporaries per scene): warm up over them
        res["reference_numpy_data_path_dataset_getitem"] = per_call_ms(lambda i: ds[i % 16], 48, warmup=48)
    res = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}
    text = json.dumps(res, indent=1)
    print(text)
    if args.json:
        with open(args.json, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
