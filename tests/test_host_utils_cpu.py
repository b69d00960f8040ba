"""Host helpers between the network and the result files (kitti_utils, calibration, bbox_transform, object3d mirrors)
against golden vectors produced by the REFERENCE modules (tools/make_host_utils_fixture.py): box corners, image
projections, bin-based decoding in the RPN / RCNN configurations, label parsing -- bit for bit.  Where the reference
tree exists (build container) the live modules are also run on fresh random inputs."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
fx = importlib.import_module("make_host_utils_fixture")


def _mine():
    return load("kitti_utils"), load("calibration"), load("bbox_transform")


def test_mirrors_equal_reference_golden(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "host_utils.npz"))
    ku, cal, bt = _mine()
    boxes, pts, dec = fx.inputs(0)
    got = fx.evaluate(ku, cal, bt, fx.calib_file(str(tmp_path)), boxes, pts, dec, torch.tensor(fx.MEAN_SIZE))
    for k, v in got.items():
        assert v.dtype == z[k].dtype and np.array_equal(v, z[k], equal_nan=True), k
    lf = tmp_path / "label.txt"
    lf.write_text("\n".join(fx.LABEL_LINES) + "\n")
    objs = ku.get_objects_from_label(str(lf))
    assert np.array_equal(ku.objs_to_boxes3d(objs), z["obj_boxes3d"])
    assert [o.level for o in objs] == list(z["obj_level"])
    assert [o.to_kitti_format() for o in objs] == list(z["obj_text"])


@pytest.mark.skipif(not os.path.isdir(fx.REF), reason="reference tree not present")
def test_mirrors_equal_live_reference_other_seeds(tmp_path):
    rku, rcal, rbt = fx.reference_modules()
    ku, cal, bt = _mine()
    path = fx.calib_file(str(tmp_path))
    for seed in (1, 2, 3):
        boxes, pts, dec = fx.inputs(seed, n_boxes=200, n_pts=4000, n_rows=3000)
        want = fx.evaluate(rku, rcal, rbt, path, boxes, pts, dec, fx.OnCpu(torch.tensor(fx.MEAN_SIZE)))
        got = fx.evaluate(ku, cal, bt, path, boxes, pts, dec, torch.tensor(fx.MEAN_SIZE))
        for k in want:
            assert np.array_equal(got[k], want[k], equal_nan=True), (seed, k)
    # float64 boxes (eval_rcnn.py hands float32; the dtype promotion must still be the reference's)
    b64 = fx.inputs(4)[0].astype(np.float64)
    assert np.array_equal(ku.boxes3d_to_corners3d(b64), rku.boxes3d_to_corners3d(b64))


@pytest.mark.skipif(not os.path.isdir(fx.REF), reason="reference tree not present")
def test_calibration_corners_of_the_interface(tmp_path):
    """the parts of Calibration the fixture does not exercise: attributes, img_to_rect, float64 inputs, zero depth"""
    _, rcal, _ = fx.reference_modules()
    cal = load("calibration")
    path = fx.calib_file(str(tmp_path))
    a, b = rcal.Calibration(path), cal.Calibration(path)
    for k in ("P2", "R0", "V2C", "cu", "cv", "fu", "fv", "tx", "ty"):
        x, y = np.asarray(getattr(a, k)), np.asarray(getattr(b, k))
        assert x.dtype == y.dtype and np.array_equal(x, y), k
    raw_a, raw_b = rcal.get_calib_from_file(path), cal.get_calib_from_file(path)
    assert set(raw_a) == set(raw_b) and all(np.array_equal(raw_a[k], raw_b[k]) for k in raw_a)
    rs = np.random.RandomState(0)
    u, v, d = (rs.uniform(0, s, 100).astype(np.float32) for s in (1242, 375, 70))
    assert np.array_equal(a.img_to_rect(u, v, d), b.img_to_rect(u, v, d))
    p64 = rs.uniform(-40, 70, (1000, 3))
    assert a.lidar_to_rect(p64).dtype == b.lidar_to_rect(p64).dtype and np.array_equal(a.lidar_to_rect(p64), b.lidar_to_rect(p64))
    pz = rs.uniform(-40, 70, (50, 3)).astype(np.float32)
    pz[::5, 2] = 0                                                   # z == 0 divides by 1e-9
    (ia, da), (ib, db) = a.rect_to_img(pz.copy()), b.rect_to_img(pz.copy())
    assert np.array_equal(ia, ib, equal_nan=True) and np.array_equal(da, db)
    assert np.array_equal(a.cart_to_hom(pz), b.cart_to_hom(pz))
