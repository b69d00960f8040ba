#!/usr/bin/env python
"""Diagnostic: per backbone level, how many clouds of a bench batch pass the exact FPS prefix test (fused.fps_gather with
fps_ordered=True) and what the check + guarded launches cost.  python tools/diag_fps_prefix.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
fz = importlib.import_module(PKG + ".fused")
cabi = importlib.import_module(PKG + ".cabi")
syn = importlib.import_module(PKG + ".synthetic")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    B = 16
    xyz = torch.from_numpy(syn.make_clouds("lidar", B, 16384, seed=1234)).cuda()
    cur = xyz
    for npoint in (4096, 1024, 256, 64):
        N = cur.shape[1]
        if N == 16384:
            _, nxt = fz.fps_gather(cur, npoint)
            print("level 16384->4096: %.3f ms" % timed(lambda: fz.fps_gather(cur, npoint)))
            cur = nxt
            continue
        viol = torch.zeros((B,), dtype=torch.int32, device="cuda")
        dmin = torch.empty((B, npoint), dtype=torch.float32, device="cuda")
        cabi.call("pn2_fps_prefix_check_f32", cabi.ptr(cur), cabi.ptr(dmin), cabi.ptr(viol), cabi.i32(B), cabi.i32(N), cabi.i32(npoint))
        idx = torch.empty((B, npoint), dtype=torch.int32, device="cuda")
        t_check = timed(lambda: cabi.call("pn2_fps_prefix_check_f32", cabi.ptr(cur), cabi.ptr(dmin), cabi.ptr(viol.clone()),
                                          cabi.i32(B), cabi.i32(N), cabi.i32(npoint)))
        t_guard = timed(lambda: cabi.call("pn2_fps_guarded_f32", cabi.ptr(cur), cabi.ptr(idx), cabi.ptr(viol), cabi.i32(B),
                                          cabi.i32(N), cabi.i32(npoint)))
        t_full = timed(lambda: cabi.call("pn2_fps_f32", cabi.ptr(cur), cabi.ptr(None), cabi.ptr(idx), cabi.i32(B), cabi.i32(N),
                                         cabi.i32(npoint)))
        print("level %d->%d: viol %s | check %.3f ms, guarded %.3f ms, full loop %.3f ms"
              % (N, npoint, viol.cpu().tolist(), t_check, t_guard, t_full))
        _, cur = fz.fps_gather(cur, npoint, fps_ordered=True)


if __name__ == "__main__":
    main()
