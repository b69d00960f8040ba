"""GPU-box diagnostic: where does the new rotated-overlap kernel differ from the reference's?"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import legacy
ic = importlib.import_module("3d_adapt_auto_driving_b200.iou3d_cuda")
g = np.load(os.path.join(ROOT, "tests/golden/iou3d_legacy.npz"))
a, b = torch.from_numpy(g["a"]).cuda(), torch.from_numpy(g["b"]).cuda()
ov = torch.zeros((300, 200), device="cuda"); ic.boxes_overlap_bev_gpu(a, b, ov)
ref = legacy.boxes_overlap_bev(a, b)
gold = torch.from_numpy(g["overlap"]).cuda()
print("legacy == golden:", torch.equal(ref, gold))
nz = (ref != 0) | (ov != 0)
d = (ov - ref).abs()
print("nonzero pairs", int(nz.sum()), "mismatching", int((ov != ref).sum()), "max abs diff", float(d.max()),
      "zero-pattern equal", torch.equal(ov == 0, ref == 0))
bad = (ov != ref).nonzero()[:10]
for i, j in bad.tolist():
    print(i, j, float(ov[i, j]), float(ref[i, j]), a[i].tolist(), b[j].tolist())
