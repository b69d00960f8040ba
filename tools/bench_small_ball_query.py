"""Ball query and FPS on the ROI-local clouds of the RCNN stage (1600 clouds of 512 / 128 points): the culled kernel against the
brute-force kernel, same lists.   python tools/bench_small_ball_query.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
fz = importlib.import_module(PKG + ".fused")
cabi = importlib.import_module(PKG + ".cabi")
inf = importlib.import_module(PKG + ".inference")
syn = importlib.import_module(PKG + ".synthetic")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    torch.manual_seed(0)
    model = inf.build_model(seed=0, device="cuda")
    pts = torch.from_numpy(syn.make_clouds("lidar", 16, 16384, seed=1024)).cuda()
    with torch.no_grad():
        out = dict(model.rpn_stage({"pts_input": pts}))
        _, rcnn_in = model.proposal_stage(out)
        pooled = model.rcnn_net._pool_rois_canonical(rcnn_in)            # (1600, 512, 136)
    xyz = pooled[..., 0:3].contiguous()
    for M, radius, ns in ((128, 0.2, 64), (32, 0.4, 64)):
        B, N, _ = xyz.shape
        t_fps = timed(lambda: fz.fps_gather(xyz, M))
        _, centres = fz.fps_gather(xyz, M)
        res = {}
        for tag, use_order in (("culled", True), ("brute force", False)):
            idx = torch.empty((B, M, ns), dtype=torch.int32, device="cuda")
            order = torch.empty((B, M), dtype=torch.int32, device="cuda") if use_order else None
            fn = lambda: cabi.call("pn2_ball_query_culled_fill_f32", cabi.ptr(centres), cabi.ptr(xyz), cabi.ptr(idx), cabi.ptr(None),
                                   cabi.ptr(order), cabi.i32(B), cabi.i32(N), cabi.i32(M), cabi.f32(radius), cabi.i32(ns),
                                   cabi.f32(0.0), cabi.i32(0), cabi.ptr(None), cabi.ptr(None))
            ms = timed(fn)
            res[tag] = idx.clone()
            print("%d clouds of %d points, %d centres, r = %.1f, nsample %d: %-11s %.3f ms" % (B, N, M, radius, ns, tag, ms), flush=True)
        print("    same lists: %s; FPS %d -> %d: %.3f ms" % (bool(torch.equal(res["culled"], res["brute force"])), N, M, t_fps), flush=True)
        xyz = centres


if __name__ == "__main__":
    main()
