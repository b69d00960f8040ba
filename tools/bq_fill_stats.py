"""How full are the ball-query groups on the benchmark batch?  A group with cnt < nsample hits is padded with copies of
its first hit (ball_query_gpu.cu:35-39), and duplicated rows cannot change a max-pool -- the fraction of UNIQUE rows is
the work a duplicate-skipping SA kernel would do.  python tools/bq_fill_stats.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"
fz = importlib.import_module(PKG + ".fused")
inf = importlib.import_module(PKG + ".inference")
syn = importlib.import_module(PKG + ".synthetic")
dev = torch.device("cuda:0")
model = inf.build_model(seed=0, device=dev)
stats = []
def wrap(name, fn, idx_pos):
    def f(*a, **k):
        idx = a[idx_pos]
        cnt = 1 + (idx[..., 1:] != idx[..., :1]).sum(-1)
        stats.append((name, tuple(idx.shape), float(cnt.float().mean()), float((cnt == idx.shape[-1]).float().mean()),
                      float(cnt.sum()) / idx.numel()))
        return fn(*a, **k)
    return f
fz.sa_fused_tc = wrap("sa_fused", fz.sa_fused_tc, 1)
fz.sa_group_linear = wrap("sa_group_linear", fz.sa_group_linear, 1)
pts = torch.from_numpy(syn.make_clouds("lidar", 16, 16384, seed=1024)).to(dev)
with torch.no_grad():
    model({"pts_input": pts})
for s in stats:
    print("%-16s idx %-22s mean unique %6.2f  full groups %5.1f%%  unique rows / rows %.3f" % (s[0], s[1], s[2], 100 * s[3], s[4]))
