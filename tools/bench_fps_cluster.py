"""FPS latency AND SM-time (latency x CTAs, the figure that matters once several batches are in flight) for every kernel
variant: the cluster kernel of csrc/fps.cu (1 / 2 / 4 / 8 CTAs per cloud, every point updated every round) and the pruned
one-CTA kernel of csrc/fps_cells.cu (4 / 8 / 16 warps).   python tools/bench_fps_cluster.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
cabi = importlib.import_module(PKG + ".cabi")
syn = importlib.import_module(PKG + ".synthetic")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for kind in ("lidar", "uniform"):
    for B, N, M in ((16, 16384, 4096), (8, 16384, 4096), (16, 4096, 1024), (16, 8192, 2048)):
        if kind == "uniform" and (B, N) != (16, 16384):
            continue
        xyz = torch.from_numpy(syn.make_clouds(kind, B, N, seed=1024)).cuda()
        ref = None
        variants = [("cluster", c) for c in ((4, 2, 8) if N > 4096 else (1, 2))] + [("cells", w) for w in (8, 4, 16)]
        for what, arg in variants:
            idx = torch.empty((B, M), dtype=torch.int32, device="cuda")
            name = "pn2_fps_cluster_f32" if what == "cluster" else "pn2_fps_cells_f32"
            fn = lambda: cabi.call(name, cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(B), cabi.i32(N), cabi.i32(M),
                                   cabi.i32(arg))
            ms = timed(fn)
            ctas = B * (arg if what == "cluster" else 1)
            if ref is None:
                ref = idx.clone()
            scan_gb = 16.0 * B * (M - 1) * N / 1e9
            print("%-7s B=%d %d->%d %s %2d: %.3f ms (%4.0f cycles/round at 1.965 GHz), %3d CTAs -> %6.1f SM-ms; %.2f TB/s of scan bytes; "
                  "same indices: %s" % (kind, B, N, M, what, arg, ms, ms * 1e-3 * 1.965e9 / (M - 1), ctas, ms * ctas,
                                        scan_gb / ms, bool(torch.equal(idx, ref))), flush=True)
