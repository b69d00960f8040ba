"""In-kernel stopwatch of the transposed fused SA kernel (csrc/sa_fused_t_tc.cu) at the RCNN SA1 / SA2 shapes of the
benchmark: ROI clouds of 512 pooled points with the bench's duplicate structure (a real detector step produces them),
dense and duplicate-skipping mode.   python tools/prof_sat.py"""
import ctypes
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
p2u = importlib.import_module(PKG + ".pointnet2_utils")

names = {0: "MMA warp total", 1: "MMA wait operand ring", 2: "MMA wait acc2 free (E2)", 3: "MMA wait A2 (E2)",
         4: "MMA wait acc3 free (E3)", 8: "EPI wait acc2 full (M2)", 9: "EPI wait A2 free (M3)", 10: "EPI wait acc3 full (M3)",
         12: "EPI E2 work", 13: "EPI E3 work", 14: "PROD wait row metadata", 15: "PROD wait free stage", 16: "PROD total"}


def run(tag, h, idx, xyz, centres, wxyz, l2, l3, out, dbg=0):
    for _ in range(2):
        fz.sa_fused_tc(h, idx, xyz, centres, wxyz, l2, l3, out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fz.sa_fused_tc(h, idx, xyz, centres, wxyz, l2, l3, out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    ncta = 148 * 4
    prof = torch.zeros((ncta * 32,), dtype=torch.int64, device="cuda")
    cabi.lib().pn2_sa_fused_t_set_profile(ctypes.c_void_p(prof.data_ptr()))
    cabi.lib().pn2_sa_fused_t_set_debug(dbg)
    fz.sa_fused_tc(h, idx, xyz, centres, wxyz, l2, l3, out)
    torch.cuda.synchronize()
    cabi.lib().pn2_sa_fused_t_set_debug(0)
    cabi.lib().pn2_sa_fused_t_set_profile(ctypes.c_void_p(0))
    pr = prof.view(ncta, 32).double().cpu()
    pr = pr[pr[:, 5] > 0]
    tiles = pr[:, 5]
    print("%s: %.3f ms (incl. compaction launches); %d CTAs with work, %.1f tiles each; cycles per tile:" % (tag, ms, pr.shape[0], float(tiles.mean())))
    for k, n in names.items():
        print("    %-28s %9.0f" % (n, float((pr[:, k] / tiles).mean())))


def main():
    torch.manual_seed(0)
    model = inf.build_model(seed=0, device="cuda")
    pts = torch.from_numpy(syn.make_clouds("lidar", 16, 16384, seed=1024)).cuda()
    with torch.no_grad():
        out = dict(model.rpn_stage({"pts_input": pts}))
        _, rcnn_in = model.proposal_stage(out)
        pooled = model.rcnn_net._pool_rois_canonical(rcnn_in)            # (1600, 512, 136)
    xyz = pooled[..., 0:3].contiguous()
    R, S, _ = xyz.shape
    g = torch.Generator(device="cpu").manual_seed(1)
    for M, radius, ns, c3, tag in ((128, 0.2, 64, 128, "RCNN SA1 [128,128,128]"), (32, 0.4, 64, 256, "RCNN SA2 [128,128,256]")):
        if M == 32:
            _, xyz = fz.fps_gather(xyz, 128)
            S = 128
        _, centres = fz.fps_gather(xyz, M)
        idx = p2u.ball_query(radius, ns, xyz, centres)
        C = 128
        h = torch.randn((R * S, C), device="cuda")
        wxyz = torch.randn((3, C), device="cuda")
        l2 = fz.PackedLayer((torch.randn((C, C), generator=g) / C ** 0.5).cuda(), torch.randn(C, generator=g).cuda(), True)
        l3 = fz.PackedLayer((torch.randn((c3, C), generator=g) / C ** 0.5).cuda(), torch.randn(c3, generator=g).cuda(), True)
        out_t = torch.empty((R * M, c3), device="cuda")
        uniq = int((idx[:, :, 1:] != idx[:, :, :1]).sum().item()) + R * M
        print("%s: %d grouped rows, %d unique (%.1f %%), %d after 8-row alignment" % (
            tag, R * M * ns, uniq, 100.0 * uniq / (R * M * ns), int(fz.group_compact(idx, align=8)[2].item())))
        for skip, align, dbg in ((True, 8, 0), (True, 8, 1), (True, 1, 0), (False, 8, 0)):
            fz.SA_SKIP_DUPLICATES, fz.SA_COMPACT_ALIGN = skip, align
            run("%s, %s%s" % (tag, ("duplicate-skipping, align %d" % align) if skip else "dense", ", WHAT-IF free pooling epilogue (stopwatch pass only)" if dbg else ""),
                h, idx, xyz, centres, wxyz, l2, l3, out_t, dbg)
    fz.SA_SKIP_DUPLICATES, fz.SA_COMPACT_ALIGN = True, 8


if __name__ == "__main__":
    main()
