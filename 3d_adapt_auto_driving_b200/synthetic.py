"""Seeded synthetic KITTI-shaped inputs; no dataset is available offline.

Clouds are (N,3) float32 in rect-camera coordinates inside PC_AREA_SCOPE
x in [-40,40], y in [-1,3], z in [0,70.4] (pointrcnn/lib/config.py:28-30).
"""
import numpy as np

SCOPE = ((-40.0, 40.0), (-1.0, 3.0), (0.0, 70.4))


def uniform_cloud(rng, n):
    lo = np.array([s[0] for s in SCOPE], np.float32)
    hi = np.array([s[1] for s in SCOPE], np.float32)
    return (rng.random_sample((n, 3)).astype(np.float32) * (hi - lo) + lo).astype(np.float32)


def lidar_cloud(rng, n, n_cars=6):
    """64-ring scan of a ground plane at y ~ 1.65 m plus a few car-sized boxes of points,
    clipped to the scope, then sampled/padded to exactly n points the way the reference
    dataset does (duplication, pointrcnn/lib/datasets/kitti_rcnn_dataset.py:310-320)."""
    rings = 64
    az = rng.uniform(-np.pi / 4 * 1.2, np.pi / 4 * 1.2, size=(n * 3,)).astype(np.float32)
    pitch = np.deg2rad(rng.randint(0, rings, size=az.shape) * (26.8 / rings) + 2.0).astype(np.float32)
    h = 1.65 + rng.normal(0, 0.02, size=az.shape).astype(np.float32)
    rng_xy = h / np.tan(pitch)
    z = rng_xy * np.cos(az)
    x = rng_xy * np.sin(az)
    y = np.full_like(x, 1.65) + rng.normal(0, 0.03, size=x.shape).astype(np.float32)
    pts = [np.stack([x, y, z], 1)]
    for _ in range(n_cars):
        c = np.array([rng.uniform(-15, 15), 0.9, rng.uniform(6, 45)], np.float32)
        dims = np.array([1.6, 1.5, 3.9], np.float32)  # w(x) h(y) l(z) before rotation
        p = (rng.random_sample((600, 3)).astype(np.float32) - 0.5) * dims
        ry = rng.uniform(-np.pi, np.pi)
        cs, sn = np.cos(ry), np.sin(ry)
        p = np.stack([p[:, 0] * cs + p[:, 2] * sn, p[:, 1], -p[:, 0] * sn + p[:, 2] * cs], 1)
        pts.append((p + c).astype(np.float32))
    pts = np.concatenate(pts, 0).astype(np.float32)
    m = ((pts[:, 0] > SCOPE[0][0]) & (pts[:, 0] < SCOPE[0][1]) & (pts[:, 1] > SCOPE[1][0]) & (pts[:, 1] < SCOPE[1][1])
         & (pts[:, 2] > SCOPE[2][0]) & (pts[:, 2] < SCOPE[2][1]))
    pts = pts[m]
    if len(pts) >= n:
        choice = rng.choice(len(pts), n, replace=False)
    else:
        choice = np.concatenate([np.arange(len(pts)), rng.choice(len(pts), n - len(pts), replace=True)])
    return np.ascontiguousarray(pts[choice])


def tie_heavy_cloud(rng, n, unique=None):
    """unique points padded to n by duplication and shuffled: exact distance ties are routine
    in the reference's input (kitti_rcnn_dataset.py:310-320), FPS must break them the same way.
    Coordinates are additionally snapped to a 0.25 m lattice so DISTINCT points tie too."""
    unique = unique or (n * 3) // 4
    base = uniform_cloud(rng, unique)
    base = (np.round(base * 4) / 4).astype(np.float32)
    extra = base[rng.choice(unique, n - unique, replace=True)]
    pts = np.concatenate([base, extra], 0)
    rng.shuffle(pts)
    return np.ascontiguousarray(pts)


def make_clouds(kind, b, n, seed):
    rng = np.random.RandomState(seed)
    fn = {"uniform": uniform_cloud, "lidar": lidar_cloud, "ties": tie_heavy_cloud}[kind]
    return np.stack([fn(rng, n) for _ in range(b)], 0)
