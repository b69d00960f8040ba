"""Small driver for ncu: the RCNN SA1 shared-MLP layers in isolation (gather layer 2, pooled layer 3)
at a quarter of the batch-16 size.  python tools/prof_tc.py [iters]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
fz = importlib.import_module(PKG + ".fused")
p2u = importlib.import_module(PKG + ".pointnet2_utils")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
R, S, M, ns, C = 400, 512, 128, 64, 128
xyz = (torch.rand((R, S, 3), device="cuda") - 0.5) * torch.tensor([4.0, 2.0, 6.0], device="cuda")
idx_c, centres = fz.fps_gather(xyz, M)
idx = p2u.ball_query(0.2, ns, xyz, centres)
h = torch.randn((R * S, C), device="cuda")
wxyz = torch.randn((3, C), device="cuda")
g = torch.Generator(device="cpu").manual_seed(1)
l2w = (torch.randn((C, C), generator=g) / C ** 0.5).cuda()
l3w = (torch.randn((C, C), generator=g) / C ** 0.5).cuda()
l2 = fz.PackedLayerTC(l2w, torch.randn(C, generator=g).cuda(), True)
l3 = fz.PackedLayerTC(l3w, torch.randn(C, generator=g).cuda(), True)
out2 = torch.empty((R * M, C), device="cuda")
mid = torch.empty((R * M * ns, C), device="cuda")
out = torch.empty((R * M, C), device="cuda")
for i in range(iters):
    s, e, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    fz.sa_group_linear_tc(h, idx, xyz, centres, wxyz, l2, out=mid)
    e.record()
    fz.linear_tc(mid, l3, out=out, pool=ns)
    e2.record()
    torch.cuda.synchronize()
    e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L2, L3 = fz.PackedLayer(l2w, l2.b, True), fz.PackedLayer(l3w, l3.b, True)
    fz.SA_TRANSPOSED = False
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out2)
    e2.record()
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out2)
    e3.record()
    fz.SA_TRANSPOSED = True
    out3 = torch.full_like(out2, -7.0)
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out3)
    e3b = torch.cuda.Event(enable_timing=True)
    e3b.record()
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out3)
    e4.record()
    torch.cuda.synchronize()
    print("gather layer %.3f ms   pooled layer %.3f ms   fused SA %.3f ms   fused SA (transposed L3) %.3f ms  (rows %d)  "
          "fused==layered: %.2e  transposed==fused: %.2e" % (
        s.elapsed_time(e), e.elapsed_time(e2), e2.elapsed_time(e3), e3b.elapsed_time(e4), R * M * ns,
        float((out - out2).abs().max() / out.abs().max()), float((out3 - out2).abs().max() / out2.abs().max())))
fz.SA_TRANSPOSED = False

# ---- in-kernel stopwatch of the fused SA kernel (cycles per role blocked on each barrier) ----
import ctypes
cabi = importlib.import_module(PKG + ".cabi")
NCTA = 148 * 4    # persistent_grid() launches up to 4 CTAs per SM
names = {0: "MMA thread total", 1: "MMA wait operand ring", 2: "MMA wait acc2 empty (E2)", 3: "MMA wait A2 full (E2)",
         4: "MMA wait acc3 empty (E3)", 6: "MMA issue layer 2 (24 MMAs)", 7: "MMA issue layer 3 (24 MMAs)", 8: "EPI wait acc2 full (M2)", 9: "EPI wait A2 empty (M3)", 10: "EPI wait acc3 full (M3)",
         12: "EPI E2 work", 13: "EPI E3 work", 14: "PROD wait row metadata", 15: "PROD wait free stage",
         16: "PROD total", 17: "META wait free slot", 18: "META total"}
L2, L3 = fz.PackedLayer(l2w, l2.b, True), fz.PackedLayer(l3w, l3.b, True)
cols = {}
for dbg in (0, 1, 2, 3):      # bit0: E3 without its TMEM loads / pooling, bit1: E2 without its TMEM ld/st
    prof = torch.zeros((NCTA * 32,), dtype=torch.int64, device="cuda")
    cabi.lib().pn2_sa_fused_tc_set_debug(dbg)
    cabi.lib().pn2_sa_fused_tc_set_profile(ctypes.c_void_p(prof.data_ptr()))
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, L2, L3, out2)
    torch.cuda.synchronize()
    cabi.lib().pn2_sa_fused_tc_set_profile(ctypes.c_void_p(0))
    cabi.lib().pn2_sa_fused_tc_set_debug(0)
    pr = prof.view(NCTA, 32).double().cpu()
    pr = pr[pr[:, 5] > 0]
    tiles = pr[:, 5].clamp_min(1)
    cols[dbg] = {k: float((pr[:, k] / tiles).mean()) for k in names}
print("fused SA kernel, cycles per tile (mean over %d CTAs, %.1f tiles each); columns: product | no E3 | no E2 | neither" % (
    pr.shape[0], float(tiles.mean())))
for k, n in names.items():
    print("  %-28s %9.0f %9.0f %9.0f %9.0f" % (n, cols[0][k], cols[1][k], cols[2][k], cols[3][k]))
