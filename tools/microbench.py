"""Per-op device timings (CUDA events, warm-up 3, median of 10): new sm_100a kernels next to
the reference's own kernels (oracle/_ref).  Writes JSON lines to the given file.
Inputs are far larger than... NOT larger than L2 for these ops (a cloud is 196 KB); the ops
are on-chip scans, so L2 flushing between iterations is done explicitly (256 MB memset)."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
p2u = importlib.import_module(PKG + ".pointnet2_utils")
cabi = importlib.import_module(PKG + ".cabi")
syn = importlib.import_module(PKG + ".synthetic")
from oracle import legacy

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main(out_path):
    rows = []

    def rec(name, cfg, new_ms, leg_ms, alg_bytes=None):
        r = {"op": name, "cfg": cfg, "new_ms": new_ms[0], "new_best_ms": new_ms[1],
             "legacy_ms": leg_ms[0] if leg_ms else None}
        if alg_bytes:
            r["scan_GBps"] = alg_bytes / (new_ms[0] * 1e-3) / 1e9
            r["frac_of_hbm_6548"] = r["scan_GBps"] / 6548.5
        rows.append(r)
        print(json.dumps(r), flush=True)

    have_legacy = legacy.available()
    for kind in ("lidar",):
        for b, n, m in [(8, 16384, 4096), (16, 16384, 4096), (16, 4096, 1024), (16, 1024, 256), (1600, 512, 128), (1600, 128, 32)]:
            xyz = torch.from_numpy(syn.make_clouds(kind, min(b, 16), n, seed=1024)).to(dev)
            if b > 16:
                xyz = xyz.repeat(b // 16, 1, 1).contiguous()
            temp = torch.empty((b, n), device=dev)
            idx = torch.empty((b, m), dtype=torch.int32, device=dev)

            cur = [0]

            def run_new():
                cabi.call("pn2_fps_cluster_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(b), cabi.i32(n),
                          cabi.i32(m), cabi.i32(cur[0]))
            clusters = [0] if n < 2048 else [0, 1, 2, 4, 8]
            for c in clusters:
                if c and b * c > 148 * 4:
                    continue
                cur[0] = c
                t_new = timeit(run_new)
                t_leg = None
                if have_legacy and c == 0:
                    t_leg = timeit(lambda: (temp.fill_(1e10), legacy.fps(xyz, m, temp)))
                rec("fps", {"b": b, "n": n, "m": m, "cluster": c, "kind": kind}, t_new, t_leg, b * (m - 1) * n * 16)

    xyz = torch.from_numpy(syn.make_clouds("lidar", 16, 16384, seed=1024)).to(dev)
    idx = p2u.furthest_point_sample(xyz, 4096)
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for b in (8, 16):
        for r, ns in [(0.1, 16), (0.5, 32), (0.1, 64)]:
            x, nx = xyz[:b].contiguous(), new_xyz[:b].contiguous()
            t_new = timeit(lambda: p2u.ball_query(r, ns, x, nx))
            t_leg = timeit(lambda: legacy.ball_query(r, ns, x, nx)) if have_legacy else None
            rec("ball_query", {"b": b, "n": 16384, "m": 4096, "r": r, "ns": ns}, t_new, t_leg, b * 4096 * 16384 * 12)
        i0 = torch.zeros((b, 4096, 16), dtype=torch.int32, device=dev)
        i1 = torch.zeros((b, 4096, 32), dtype=torch.int32, device=dev)
        x, nx = xyz[:b].contiguous(), new_xyz[:b].contiguous()
        t_new = timeit(lambda: cabi.call("pn2_ball_query_dual_f32", cabi.ptr(nx), cabi.ptr(x), cabi.ptr(i0), cabi.ptr(i1),
                                         cabi.i32(b), cabi.i32(16384), cabi.i32(4096), cabi.f32(0.1), cabi.i32(16),
                                         cabi.f32(0.5), cabi.i32(32)))
        rec("ball_query_dual", {"b": b, "n": 16384, "m": 4096, "r": [0.1, 0.5], "ns": [16, 32]}, t_new, None, b * 4096 * 16384 * 12)

    b = 16
    t_new = timeit(lambda: p2u.three_nn(xyz, new_xyz))
    t_leg = timeit(lambda: legacy.three_nn(xyz, new_xyz)) if have_legacy else None
    rec("three_nn", {"b": b, "n": 16384, "m": 4096}, t_new, t_leg, b * 4096 * 16384 * 12)
    dist, i3 = p2u.three_nn(xyz, new_xyz)
    feats = torch.randn((b, 256, 4096), device=dev)
    w = torch.rand((b, 16384, 3), device=dev)
    t_new = timeit(lambda: p2u.three_interpolate(feats, i3, w))
    t_leg = timeit(lambda: legacy.three_interpolate(feats, i3, w)) if have_legacy else None
    rec("three_interpolate", {"b": b, "c": 256, "m": 4096, "n": 16384}, t_new, t_leg, b * 256 * (4096 + 16384) * 4 + b * 16384 * 24)
    bq = p2u.ball_query(0.5, 32, xyz, new_xyz)
    f96 = torch.randn((b, 96, 16384), device=dev)
    t_new = timeit(lambda: p2u.grouping_operation(f96, bq))
    t_leg = timeit(lambda: legacy.group(f96, bq)) if have_legacy else None
    rec("group_points", {"b": b, "c": 96, "n": 16384, "m": 4096, "ns": 32}, t_new, t_leg, b * 96 * (16384 + 4096 * 32) * 4 + b * 4096 * 32 * 4)
    with open(out_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/microbench.jsonl")
