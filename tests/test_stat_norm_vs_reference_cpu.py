"""stat_norm mirror against the REFERENCE stat_norm/norm.py imported live (tools/make_statnorm_fixture.load_reference:
HOME redirected because its config_path import creates ~/scratch/..., np.ones patched for the uint8 occupancy map that
overflows on NumPy 2) on RANDOM scenes: several label sets, box poses, point densities and all four
(avoid_conflict, align_front) combinations -- rescaled points, ratios and regenerated label text bit-identical.
Build container only; tests/test_stat_norm.py carries the committed goldens for one scene."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
fx = importlib.import_module("make_statnorm_fixture")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(fx.REF, "stat_norm", "norm.py")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def reference():
    home, path = os.environ.get("HOME"), list(sys.path)
    mods = set(sys.modules)
    try:
        ref = fx.load_reference()
        from utils.kitti_util import Calibration as RefCalibration
        from utils.object_3d import Object3d as RefObject3d
    finally:
        if home is not None:
            os.environ["HOME"] = home
        sys.path[:] = path
        for k in set(sys.modules) - mods:
            if k == "utils" or k.startswith("utils.") or k == "config_path":
                del sys.modules[k]
    return ref, RefCalibration, RefObject3d


def _random_labels(rs, n):
    lines = []
    for _ in range(n):
        cls = rs.choice(["Car", "Van", "Pedestrian", "Cyclist", "Truck"], p=[0.5, 0.2, 0.1, 0.1, 0.1])
        h, w, l = rs.uniform(1.3, 2.2), rs.uniform(1.4, 2.0), rs.uniform(3.2, 5.5)
        x, y, z = rs.uniform(-20, 20), rs.uniform(1.2, 1.9), rs.uniform(6, 60)
        ry = rs.uniform(-np.pi, np.pi)
        alpha = rs.uniform(-np.pi, np.pi)
        lines.append("%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f" % (
            cls, rs.uniform(0, 0.5), rs.randint(0, 3), alpha, 100, 100, 200, 200, h, w, l, x, y, z, ry))
    return lines


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scenes_bit_identical(reference, tmp_path, seed):
    ref, RefCalibration, RefObject3d = reference
    norm, ku, o3 = load("stat_norm.norm"), load("stat_norm.kitti_util"), load("stat_norm.object_3d")
    cpath = tmp_path / "000000.txt"
    cpath.write_text(fx.CALIB_TXT)
    rs = np.random.RandomState(seed)
    lines = _random_labels(rs, int(rs.randint(3, 9)))
    calib_r, calib_m = RefCalibration(str(cpath)), ku.Calibration(str(cpath))
    labels_r, labels_m = [RefObject3d(l) for l in lines], [o3.Object3d(l) for l in lines]
    velo = fx.make_scene(calib_r, labels_r)                      # points inside the first boxes + an environment wall
    velo[2000:, :3] += rs.normal(0, 0.05, (velo.shape[0] - 2000, 3)).astype(np.float32)
    map_r = ref.get_scale_map(ref.germany_car_stats, ref.us_car_stats)
    map_m = norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)
    for ac in (False, True):
        for af in (False, True):
            pts_r, ratios_r = ref.rescale_ptc(map_r, velo, labels_r, calib_r, avoid_conflict=ac, align_front=af)
            pts_m, ratios_m = norm.rescale_ptc(map_m, velo, labels_m, calib_m, avoid_conflict=ac, align_front=af)
            assert np.array_equal(np.asarray(ratios_r, np.float64), np.asarray(ratios_m, np.float64)), (ac, af)
            assert pts_r.dtype == pts_m.dtype and np.array_equal(pts_r, pts_m), (ac, af)
            new_r = ref.scale_labels(labels_r, map_r, ratios_r, calib_r, 1242, 375, align_front=af)
            new_m = norm.scale_labels(labels_m, map_m, ratios_m, calib_m, 1242, 375, align_front=af)
            assert [o.to_kitti_format() for o in new_r] == [o.to_kitti_format() for o in new_m], (ac, af)
