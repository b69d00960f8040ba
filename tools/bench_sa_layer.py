"""BASELINE.json configs[2]: one pointnet2 SA layer -- FPS 16384 -> 4096 + ball_query r = 0.1, nsample = 64 +
3-layer shared MLP [3 -> 16 -> 16 -> 32] + max-pool -- batch 8, one B200.  Reports the time of every stage of this
package's path and of the reference's own kernels (oracle/_ref, same inputs), and the FPS / ball-query scan
bandwidth against the measured HBM peak (north_star: >= 60 % on FPS + ball_query).

    python tools/bench_sa_layer.py [out.json]
"""
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
p2m = importlib.import_module(PKG + ".pointnet2_modules")
fz = importlib.import_module(PKG + ".fused")
syn = importlib.import_module(PKG + ".synthetic")
from oracle import legacy

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                                     # L2 flush between iterations (a cloud is 196 KB)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main(out_path):
    B, N, M, R, NS = 8, 16384, 4096, 0.1, 64
    out = {"config": {"batch": B, "npoints": N, "npoint": M, "radius": R, "nsample": NS, "mlp": [3, 16, 16, 32]}, "kinds": {}}
    torch.manual_seed(0)
    sa = p2m.PointnetSAModuleMSG(npoint=M, radii=[R], nsamples=[NS], mlps=[[0, 16, 16, 32]], use_xyz=True, bn=True).to(dev).eval()
    for kind in ("lidar", "uniform", "ties"):
        xyz = torch.from_numpy(syn.make_clouds(kind, B, N, seed=1024)).to(dev)
        idx = p2u.furthest_point_sample(xyz, M)
        new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        r = {}
        r["fps_ms"] = timeit(lambda: p2u.furthest_point_sample(xyz, M))
        r["ball_query_ms"] = timeit(lambda: p2u.ball_query(R, NS, xyz, new_xyz))
        with torch.no_grad():
            r["sa_layer_ms"] = timeit(lambda: sa(xyz, None))          # FPS + gather + ball query + fused MLP + max-pool
        r["mlp_pool_ms"] = max(r["sa_layer_ms"] - r["fps_ms"] - r["ball_query_ms"], 0.0)
        fps_bytes, bq_bytes = 16.0 * B * (M - 1) * N, 12.0 * B * M * N
        r["fps_scan_GBps"] = fps_bytes / r["fps_ms"] / 1e6
        r["ball_query_scan_GBps"] = bq_bytes / r["ball_query_ms"] / 1e6
        r["fps_plus_bq_frac_of_hbm"] = (fps_bytes + bq_bytes) / (r["fps_ms"] + r["ball_query_ms"]) / 1e6 / peaks["hbm_gbs"]
        if legacy.available():
            temp = torch.empty((B, N), device=dev)
            r["legacy_fps_ms"] = timeit(lambda: (temp.fill_(1e10), legacy.fps(xyz, M, temp)))
            r["legacy_ball_query_ms"] = timeit(lambda: legacy.ball_query(R, NS, xyz, new_xyz))
        out["kinds"][kind] = {k: round(v, 4) for k, v in r.items()}
        print(kind, json.dumps(out["kinds"][kind]), flush=True)
    if legacy.available():
        # the reference's whole layer: its kernels under the op-by-op module composition (cuDNN 1x1 convs); install()
        # re-points the package's extension shims at oracle/_ref for the rest of the process, so this comes last
        legacy.install(PKG)
        for m in sa.modules():
            if hasattr(m, "fused"):
                m.fused = False
        for kind in out["kinds"]:
            xyz = torch.from_numpy(syn.make_clouds(kind, B, N, seed=1024)).to(dev)
            with torch.no_grad():
                out["kinds"][kind]["legacy_sa_layer_ms"] = round(timeit(lambda: sa(xyz, None), iters=5), 4)
            print(kind, "legacy_sa_layer_ms", out["kinds"][kind]["legacy_sa_layer_ms"], flush=True)
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sa_layer.json")
