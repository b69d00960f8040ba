"""Golden vectors for stat_norm from the REFERENCE module itself (runs only where /root/reference is
mounted; the resulting tests/golden/stat_norm.npz is committed).

The reference stat_norm/norm.py is imported unmodified with (1) HOME redirected, because its
config_path import creates ~/scratch/driving_datasets, and (2) np.ones patched inside that module
only for the uint8 occupancy map of `postprocessing`, which overflows on NumPy 2 (norm.py:134);
int16 is what NumPy 1.x value-based casting produced.  Scene = SURVEY.md 8(d) config 1."""
import hashlib
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PN2_REFERENCE_ROOT", "/root/reference")

CALIB_TXT = """P0: 7.215377e+02 0.0 6.095593e+02 0.0 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P1: 7.215377e+02 0.0 6.095593e+02 -3.875744e+02 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P2: 7.215377e+02 0.0 6.095593e+02 4.485728e+01 0.0 7.215377e+02 1.728540e+02 2.163791e-01 0.0 0.0 1.0 2.745884e-03
P3: 7.215377e+02 0.0 6.095593e+02 -3.395242e+02 0.0 7.215377e+02 1.728540e+02 2.199936e+00 0.0 0.0 1.0 2.729905e-03
R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01
Tr_velo_to_cam: 7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 1.480755e-02 -2.717806e-01
Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01
"""

LABELS = [
    "Car 0.00 0 1.20 600.00 150.00 700.00 220.00 1.50 1.60 3.90 2.00 1.60 12.00 0.30",
    "Van 0.10 1 -1.90 300.00 140.00 420.00 230.00 2.10 1.90 5.10 -4.50 1.70 18.00 -1.40",
    "Pedestrian 0.00 0 0.40 800.00 150.00 830.00 230.00 1.75 0.60 0.80 6.00 1.60 9.00 0.10",
    "Car 0.00 2 2.80 100.00 160.00 180.00 200.00 1.45 1.55 3.60 -12.00 1.80 30.00 2.90",
    "Car 0.00 0 0.00 500.00 170.00 520.00 180.00 1.50 1.60 4.00 40.00 1.50 60.00 0.00",
]


def make_scene(calib, labels):
    """16384 velodyne points: uniform background + points placed inside the first four boxes."""
    rng = np.random.RandomState(0)
    n = 16384
    velo = np.stack([rng.uniform(0, 70, n), rng.uniform(-40, 40, n), rng.uniform(-3, 1, n), rng.uniform(0, 1, n)], 1)
    k = 0
    for obj in labels[:4]:
        m = 200
        loc = np.stack([rng.uniform(-obj.l / 2, obj.l / 2, m) * 0.98, rng.uniform(-obj.h, 0, m) * 0.98,
                        rng.uniform(-obj.w / 2, obj.w / 2, m) * 0.98], 1)
        c, s = np.cos(obj.ry), np.sin(obj.ry)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        rect = loc @ R.T + obj.t
        velo[k:k + m, :3] = calib.project_rect_to_velo(rect)
        k += m
    # a wall of "environment" points just beyond the ends of box 0, so that avoid_conflict has to back off
    obj = labels[0]
    m = 400
    side = np.where(rng.rand(m) < 0.5, -1.0, 1.0)
    loc = np.stack([side * (obj.l / 2 + rng.uniform(0.02, 0.45, m)), rng.uniform(-obj.h, -0.6, m),
                    rng.uniform(-obj.w / 2, obj.w / 2, m) * 0.9], 1)
    c, s = np.cos(obj.ry), np.sin(obj.ry)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    velo[k:k + m, :3] = calib.project_rect_to_velo(loc @ R.T + obj.t)
    return velo.astype(np.float32)


def load_reference():
    home = tempfile.mkdtemp()
    os.environ["HOME"] = home
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_norm", os.path.join(REF, "stat_norm", "norm.py"))
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "stat_norm"))
    try:
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    proxy = types.ModuleType("np_proxy")
    proxy.__dict__.update(np.__dict__)
    proxy.ones = lambda shape, dtype=None: np.ones(shape, dtype=np.int16 if dtype == np.uint8 else dtype)
    mod.np = proxy
    return mod


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main(out_path):
    ref = load_reference()
    from utils.kitti_util import Calibration
    from utils.object_3d import Object3d
    tmp = tempfile.mkdtemp()
    cpath = os.path.join(tmp, "000000.txt")
    with open(cpath, "w") as f:
        f.write(CALIB_TXT)
    calib = Calibration(cpath)
    labels = [Object3d(l) for l in LABELS]
    velo = make_scene(calib, labels)
    mapping = ref.get_scale_map(ref.germany_car_stats, ref.us_car_stats)
    out = {"velo": velo}
    for ac in (False, True):
        for af in (False, True):
            tag = "ac%d_af%d" % (ac, af)
            pts, ratios = ref.rescale_ptc(mapping, velo, labels, calib, avoid_conflict=ac, align_front=af)
            out[tag + "_pts_sha"] = sha(pts)
            out[tag + "_pts_head"] = pts[:1000].copy()
            out[tag + "_ratios"] = np.asarray(ratios, np.float64)
            binp = os.path.join(tmp, tag + ".bin")
            ref.format_lidar_data(pts, binp)
            out[tag + "_bin_sha"] = sha(np.fromfile(binp, np.uint8))
            new_labels = ref.scale_labels(labels, mapping, ratios, calib, 1242, 375, align_front=af)
            out[tag + "_labels"] = np.array("\n".join(o.to_kitti_format() for o in new_labels))
            print(tag, "ratios", ratios, "pts", pts.shape, pts.dtype)
    np.savez_compressed(out_path, calib=np.array(CALIB_TXT), labels=np.array("\n".join(LABELS)), **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "stat_norm.npz"))
