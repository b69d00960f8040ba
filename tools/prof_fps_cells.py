"""In-kernel stopwatch of the pruned FPS kernel (csrc/fps_cells.cu) at 16 x (16384 -> 4096): cycles per round in each
phase, separately for the warps that touch a cell in a round and for those that only wait.   python tools/prof_fps_cells.py"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
cabi = importlib.import_module(PKG + ".cabi")
syn = importlib.import_module(PKG + ".synthetic")

B, N, M = 16, 16384, 4096
kind = sys.argv[1] if len(sys.argv) > 1 else "lidar"
xyz = torch.from_numpy(syn.make_clouds(kind, B, N, seed=1024)).cuda()
for warps in (8, 4, 16):
    idx = torch.empty((B, M), dtype=torch.int32, device="cuda")
    prof = torch.zeros((B, warps, 8), dtype=torch.int64, device="cuda")
    call = lambda: cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(B), cabi.i32(N),
                             cabi.i32(M), cabi.i32(warps))
    call(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); call(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    cabi.lib().pn2_fps_cells_set_profile(ctypes.c_void_p(prof.data_ptr()))
    call(); torch.cuda.synchronize()
    cabi.lib().pn2_fps_cells_set_profile(ctypes.c_void_p(0))
    p = prof.double().cpu()
    rounds = M - 1
    nupd = p[..., 6]
    print("%s, %d warps: %.3f ms = %.0f cycles per round; per warp and round: touches a cell in %.1f %% of the rounds, "
          "%.2f cells per touching round, %.2f touched cells per round in the whole CTA" % (
              kind, warps, ms, ms * 1e-3 * 1.965e9 / rounds, 100 * float((nupd / rounds).mean()),
              float((p[..., 7].sum() / nupd.sum())), float(p[..., 7].sum() / B / rounds)))
    print("    box test (every round)            %7.0f" % float((p[..., 0] / rounds).mean()))
    print("    cell updates + records (touching) %7.0f" % float((p[..., 1] / nupd).mean()))
    print("    barrier wait, touching rounds     %7.0f" % float((p[..., 3] / nupd).mean()))
    print("    barrier wait, idle rounds         %7.0f" % float((p[..., 4] / (rounds - nupd)).mean()))
    print("    record reduce + next centre       %7.0f" % float((p[..., 5] / rounds).mean()))
