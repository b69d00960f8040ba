"""How many spatial cells does a round of furthest point sampling touch?  Numpy model of the pruning of csrc/fps_cells.cu
(Morton-ordered cells of 512 / 256 / 128 points with bounding boxes; a cell is touched when the new centre's distance to its box
is below the cell's current maximum min-distance), run before the kernel was written: ~3-5 of 32-128 cells per round on the
synthetic LiDAR and uniform clouds.   python tools/sim_fps_cells.py"""
import sys, numpy as np, importlib
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
syn = importlib.import_module('3d_adapt_auto_driving_b200.synthetic')
def morton2(ix, iz):
    def part(v):
        v = v.astype(np.uint32)
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return part(ix) | (part(iz) << 1)
def sim(kind, n=16384, m=4096, cell=512, seed=0, use3d=False):
    pts = syn.make_clouds(kind, 1, n, seed)[0].astype(np.float32)
    lo = pts.min(0); hi = pts.max(0)
    q = ((pts - lo) / (hi - lo + 1e-9) * 255).astype(np.int64)
    key = morton2(q[:,0], q[:,2])
    order = np.argsort(key, kind='stable')
    P = pts[order]
    nc = n // cell
    cells = P.reshape(nc, cell, 3)
    blo = cells.min(1); bhi = cells.max(1)
    pt = np.full(n, 1e10, np.float32)
    cmax = np.full(nc, 1e10, np.float32)
    cur = np.where(order == 0)[0][0]
    upd_hist = []
    for r in range(m - 1):
        c = P[cur]
        d = np.maximum(0, np.maximum(blo - c, c - bhi)).astype(np.float32)
        lb = (d * d).sum(1)
        need = lb < cmax
        upd_hist.append(need.sum())
        for ci in np.nonzero(need)[0]:
            sl = slice(ci * cell, (ci + 1) * cell)
            dd = ((P[sl] - c) ** 2).sum(1).astype(np.float32)
            pt[sl] = np.minimum(pt[sl], dd)
            cmax[ci] = pt[sl].max()
        cur = int(np.argmax(pt))
    u = np.array(upd_hist)
    print(kind, 'cell', cell, 'cells', nc, 'mean updated cells/round', u.mean().round(2), 'after 256:', u[256:].mean().round(2),
          'p50', np.percentile(u[256:], 50), 'p90', np.percentile(u[256:], 90), 'max', u[256:].max(), ' rounds w/ <=1:', (u[256:] <= 1).mean().round(3))
for kind in ('lidar', 'uniform'):
    for cell in (512, 256, 128):
        sim(kind, cell=cell)
