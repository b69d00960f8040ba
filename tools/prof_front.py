"""In-kernel stopwatch of the fused RCNN input chain (csrc/rcnn_front_tc.cu): cycles per tile every role spends blocked
on each barrier, plus event-timed launches at the benchmark size.   python tools/prof_front.py [rows]"""
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

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 100 * 512
g = torch.Generator(device="cpu").manual_seed(3)


def mk(cout, cin, relu):
    return fz.PackedLayer((torch.randn((cout, cin), generator=g) / cin ** 0.5).cuda(), torch.randn((cout,), generator=g).cuda(), relu)


l1, l2, lm, ls = mk(128, 5, True), mk(128, 128, True), mk(128, 256, True), mk(128, 128, False)
x = torch.randn((rows, 136), device="cuda")
out = torch.empty((rows, 128), device="cuda")
for _ in range(3):
    fz.rcnn_front(x, 8, l1, l2, lm, ls, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    fz.rcnn_front(x, 8, l1, l2, lm, ls, out=out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
gb = rows * (136 + 128) * 4 / 1e9
print("rcnn_front: %d rows, %.3f ms per launch, %.2f TB/s of compulsory bytes (%.2f GB), %.0f TFLOP/s useful"
      % (rows, ms, gb / ms, gb, 2.0 * rows * (128 * 128 * 2 + 256 * 128) / ms / 1e9))
s.record()
for _ in range(10):
    fz.linear(fz.linear_cat(fz.linear_pre(x, 5, l1, l2), x[:, 8:], lm), ls)
e.record()
torch.cuda.synchronize()
print("three separate launches: %.3f ms" % (s.elapsed_time(e) / 10))

NCTA = 148
names = {0: "MMA warp total", 1: "MMA wait operand ring", 2: "MMA wait weight ring", 3: "MMA wait EA (xyz feature, E1)",
         4: "MMA wait EA (merged, E2)", 5: "MMA wait accumulator free", 8: "EPI wait acc1 full (M1)", 9: "EPI wait acc2 full (M2X)",
         10: "EPI wait acc3 full (M3)", 11: "EPI wait EA free", 12: "EPI E1 + E2 work", 13: "EPI E3 (store H) work",
         16: "PROD total", 17: "PROD wait free stage", 18: "PROD group barrier", 19: "PROD computed step", 20: "PROD feature step (incl. load wait)",
         7: "MMA wake-up after the last EA arrival", 21: "EPI E1 + E2: tcgen05.ld + wait", 22: "EPI E1 + E2: tcgen05.st wait",
         23: "EPI E3: tcgen05.ld + wait"}
cols = {}
for mode in (0, 2):       # 2: M2F issued before M3 (the first version's order)
    cabi.lib().pn2_rcnn_front_set_mode(mode)
    for _ in range(2):
        fz.rcnn_front(x, 8, l1, l2, lm, ls, out=out)
    torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        fz.rcnn_front(x, 8, l1, l2, lm, ls, out=out)
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e) / 10
    prof = torch.zeros((NCTA * 32,), dtype=torch.int64, device="cuda")
    cabi.lib().pn2_rcnn_front_set_profile(ctypes.c_void_p(prof.data_ptr()))
    fz.rcnn_front(x, 8, l1, l2, lm, ls, out=out)
    torch.cuda.synchronize()
    cabi.lib().pn2_rcnn_front_set_profile(ctypes.c_void_p(0))
    pr = prof.view(NCTA, 32).double().cpu()
    pr = pr[pr[:, 6] > 0]
    tiles = pr[:, 6]
    cols[mode] = ({k: float((pr[:, k] / tiles).mean()) for k in names}, t)
cabi.lib().pn2_rcnn_front_set_mode(0)
print("cycles per tile (mean over %d CTAs, %.1f tiles each); columns: default order (%.3f ms) | M2F before M3 (%.3f ms)"
      % (pr.shape[0], float(tiles.mean()), cols[0][1], cols[2][1]))
for k, n in names.items():
    print("  %-34s %9.0f %9.0f" % (n, cols[0][0][k], cols[2][0][k]))
