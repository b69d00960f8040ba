"""dense vs duplicate-skipping transposed SA kernel on ball_query-shaped (padded) groups.  python tools/prof_compact.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
fz = importlib.import_module(PKG + ".fused")
torch.manual_seed(0)
R, S, M, ns, C = 400, 512, 128, 64, 128
dev = "cuda"
xyz = (torch.rand((R, S, 3), device=dev) - 0.5) * torch.tensor([4.0, 2.0, 6.0], device=dev)
centres = xyz[:, :M].contiguous()
g = torch.Generator(device="cpu").manual_seed(1)
h = torch.randn((R * S, C), device=dev)
wxyz = torch.randn((3, C), device=dev)
l2w = (torch.randn((C, C), generator=g) / C ** 0.5).to(dev)
l3w = (torch.randn((C, C), generator=g) / C ** 0.5).to(dev)
L2 = fz.PackedLayer(l2w, torch.randn(C, generator=g).to(dev), True)
L3 = fz.PackedLayer(l3w, torch.randn(C, generator=g).to(dev), True)
out = torch.empty((R * M, C), device=dev)
fz.SA_SKIP_MIN_ROWS = 0
for mean_fill in (64, 48, 32, 16, 4):
    cnt = torch.randint(max(1, 2 * mean_fill - 64), min(64, 2 * mean_fill) + 1, (R, M, 1), generator=g) if mean_fill < 64 else torch.full((R, M, 1), 64)
    perm = torch.argsort(torch.rand((R, M, S), generator=g), dim=2)[:, :, :ns].to(torch.int32)
    k = torch.arange(ns).view(1, 1, ns)
    idx = torch.where(k < cnt, perm, perm[:, :, :1]).contiguous().to(dev)
    res = {}
    for mode in (False, True):
        fz.SA_SKIP_DUPLICATES = mode
        for _ in range(2):
            fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out)
        e.record()
        torch.cuda.synchronize()
        res[mode] = s.elapsed_time(e) / 5
    # kernel-only time of the compact path (without the two compaction launches / cumsum / zero-fill)
    cm, jm, nr = fz.group_compact(idx)
    t2, (w3hi, w3lo) = L2.tc, L3.w3t
    cabi = fz.cabi
    def kern():
        cabi.call("pn2_sa_fused_t_tc_f32", fz.ptr(h), fz.i32(C), fz.ptr(idx), fz.ptr(xyz), fz.ptr(centres), fz.ptr(wxyz), fz.ptr(t2.blob),
                  fz.i32(t2.ntile), fz.i32(t2.nkb), fz.ptr(t2.b), fz.ptr(w3hi), fz.ptr(w3lo), fz.ptr(L3.b), fz.ptr(out), fz.i32(C), fz.i32(R),
                  fz.i32(S), fz.i32(M), fz.i32(ns), fz.i32(C), fz.i32(C), fz.i32(C), fz.ptr(cm), fz.ptr(jm), fz.ptr(nr), fz.i32(fz.SA_COMPACT_ALIGN))
    kern(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        kern()
    e.record(); torch.cuda.synchronize()
    print("mean fill %2d: unique rows %.3f  dense %.3f ms  compact (all) %.3f ms  compact kernel only %.3f ms" % (
        mean_fill, float(cnt.sum()) / (R * M * ns), res[False], res[True], s.elapsed_time(e) / 5))
