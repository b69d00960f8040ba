"""FPS 16384 -> 4096 at batch 16 (and 8) for every thread-block-cluster size: latency AND SM-time (latency x CTAs), the
figure that matters once several batches are in flight.   python tools/bench_fps_cluster.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
cabi = importlib.import_module(PKG + ".cabi")
syn = importlib.import_module(PKG + ".synthetic")

for B in (16, 8):
    xyz = torch.from_numpy(syn.make_clouds("lidar", B, 16384, seed=1024)).cuda()
    ref = None
    for cluster in (4, 2, 8):
        idx = torch.empty((B, 4096), dtype=torch.int32, device="cuda")
        fn = lambda: cabi.call("pn2_fps_cluster_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(B), cabi.i32(16384),
                               cabi.i32(4096), cabi.i32(cluster))
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            fn()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        if ref is None:
            ref = idx.clone()
        print("B=%d cluster %d: %.3f ms, %d CTAs -> %.1f SM-ms; same indices: %s" % (B, cluster, ms, B * cluster, ms * B * cluster,
                                                                                  bool(torch.equal(idx, ref))))
