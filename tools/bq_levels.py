"""Per-level timing of the neighbour searches on the bench clouds: brute force vs culled."""
import importlib, sys, os, torch
sys.path.insert(0, os.getcwd())
P = "3d_adapt_auto_driving_b200"
cabi, fz, syn = (importlib.import_module(P + "." + m) for m in ("cabi", "fused", "synthetic"))
from ctypes import c_void_p
dev = "cuda"
B = 16
xyz = torch.from_numpy(syn.make_clouds("lidar", B, 16384, seed=1)).to(dev)
levels = [(16384, 4096, 0.1, 16, 0.5, 32), (4096, 1024, 0.5, 16, 1.0, 32), (1024, 256, 1.0, 16, 2.0, 32), (256, 64, 2.0, 16, 4.0, 32)]
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1000
cur = xyz
pts = [xyz]
for (n, m, r0, ns0, r1, ns1) in levels:
    idx, new_xyz = fz.fps_gather(cur, m)
    i0 = torch.zeros((B, m, ns0), dtype=torch.int32, device=dev); i1 = torch.zeros((B, m, ns1), dtype=torch.int32, device=dev)
    order = torch.empty((B, m), dtype=torch.int32, device=dev)
    def brute(): cabi.call("pn2_ball_query_dual_f32", cabi.ptr(new_xyz), cabi.ptr(cur), cabi.ptr(i0), cabi.ptr(i1), cabi.i32(B), cabi.i32(n), cabi.i32(m), cabi.f32(r0), cabi.i32(ns0), cabi.f32(r1), cabi.i32(ns1))
    def culled(): cabi.call("pn2_ball_query_culled_f32", cabi.ptr(new_xyz), cabi.ptr(cur), cabi.ptr(i0), cabi.ptr(i1), cabi.ptr(order), cabi.i32(B), cabi.i32(n), cabi.i32(m), cabi.f32(r0), cabi.i32(ns0), cabi.f32(r1), cabi.i32(ns1))
    print("ball_query n=%d m=%d: brute %.1f us, culled %.1f us" % (n, m, timeit(brute), timeit(culled)), flush=True)
    cur = new_xyz; pts.append(new_xyz)
for (nb, n, m, r, ns) in ((1600, 512, 128, 0.2, 64), (1600, 128, 32, 0.4, 64)):
    small = (torch.rand((nb, n, 3), device=dev) - 0.5) * torch.tensor([2.0, 1.5, 4.5], device=dev)
    cen = small[:, :m].contiguous()
    i0 = torch.zeros((nb, m, ns), dtype=torch.int32, device=dev)
    order = torch.empty((nb, m), dtype=torch.int32, device=dev)
    def brute(): cabi.call("pn2_ball_query_f32", cabi.ptr(cen), cabi.ptr(small), cabi.ptr(i0), cabi.i32(nb), cabi.i32(n), cabi.i32(m), cabi.f32(r), cabi.i32(ns))
    def culled(): cabi.call("pn2_ball_query_culled_f32", cabi.ptr(cen), cabi.ptr(small), cabi.ptr(i0), cabi.ptr(None), cabi.ptr(order), cabi.i32(nb), cabi.i32(n), cabi.i32(m), cabi.f32(r), cabi.i32(ns), cabi.f32(0.0), cabi.i32(0))
    print("ball_query rcnn b=%d n=%d m=%d: brute %.1f us, culled %.1f us" % (nb, n, m, timeit(brute), timeit(culled)), flush=True)
for lvl in range(3, -1, -1):
    unknown, known = pts[lvl], pts[lvl + 1]
    n, m = unknown.shape[1], known.shape[1]
    d2 = torch.empty((B, n, 3), device=dev); ix = torch.empty((B, n, 3), dtype=torch.int32, device=dev)
    order = torch.empty((B, n), dtype=torch.int32, device=dev)
    def brute(): cabi.call("pn2_three_nn_f32", cabi.ptr(unknown), cabi.ptr(known), cabi.ptr(d2), cabi.ptr(ix), cabi.i32(B), cabi.i32(n), cabi.i32(m))
    def culled(): cabi.call("pn2_three_nn_culled_f32", cabi.ptr(unknown), cabi.ptr(known), cabi.ptr(d2), cabi.ptr(ix), cabi.ptr(order), cabi.i32(B), cabi.i32(n), cabi.i32(m))
    print("three_nn n=%d m=%d: brute %.1f us, culled %.1f us" % (n, m, timeit(brute), timeit(culled)), flush=True)
for m in (4096, 1024, 16384):
    order = torch.empty((B, m), dtype=torch.int32, device=dev)
    src = pts[0] if m == 16384 else pts[1 if m == 4096 else 2]
    lib = cabi.lib()
    print("fps n=16384->%d" % m if False else "", end="")
for (n, m) in ((16384, 4096), (4096, 1024), (1024, 256)):
    src = {16384: pts[0], 4096: pts[1], 1024: pts[2]}[n]
    print("fps n=%d m=%d: %.1f us" % (n, m, timeit(lambda: fz.fps_gather(src, m), it=5)), flush=True)
