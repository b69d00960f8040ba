"""stat_norm (Statistical Normalization) mirror against golden vectors produced by the REFERENCE
module itself (tools/make_statnorm_fixture.py imports /root/reference/stat_norm/norm.py unmodified and
runs rescale_ptc / format_lidar_data / scale_labels on the SURVEY 8(d) config-1 scene).
Rescaled coordinates must be bit-exact (float64 sha256 and the float32 .bin bytes)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stat_norm.npz")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    g = np.load(GOLD)
    norm = load("stat_norm.norm")
    ku = load("stat_norm.kitti_util")
    o3 = load("stat_norm.object_3d")
    d = tmp_path_factory.mktemp("statnorm")
    cpath = d / "000000.txt"
    cpath.write_text(str(g["calib"]))
    calib = ku.Calibration(str(cpath))
    labels = [o3.Object3d(l) for l in str(g["labels"]).split("\n")]
    mapping = norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)
    return g, norm, calib, labels, mapping, d


@pytest.mark.parametrize("ac", [False, True])
@pytest.mark.parametrize("af", [False, True])
def test_rescale_ptc_bit_exact_vs_reference(scene, ac, af):
    g, norm, calib, labels, mapping, d = scene
    tag = "ac%d_af%d" % (ac, af)
    pts, ratios = norm.rescale_ptc(mapping, g["velo"], labels, calib, avoid_conflict=ac, align_front=af)
    assert pts.dtype == np.float64 and pts.shape == (16384, 3)
    assert np.array_equal(np.asarray(ratios, np.float64), g[tag + "_ratios"])
    assert np.array_equal(pts[:1000], g[tag + "_pts_head"])
    assert np.array_equal(sha(pts), g[tag + "_pts_sha"])
    binp = d / (tag + ".bin")
    norm.format_lidar_data(pts, str(binp))
    raw = np.fromfile(str(binp), np.uint8)
    assert raw.size == 16384 * 4 * 4 and np.array_equal(sha(raw), g[tag + "_bin_sha"])
    assert np.all(np.fromfile(str(binp), np.float32).reshape(-1, 4)[:, 3] == 1.0)   # intensity replaced by 1.0


@pytest.mark.parametrize("ac,af", [(False, False), (True, True)])
def test_scale_labels_text_matches_reference(scene, ac, af):
    g, norm, calib, labels, mapping, d = scene
    tag = "ac%d_af%d" % (ac, af)
    new_labels = norm.scale_labels(labels, mapping, g[tag + "_ratios"].tolist(), calib, 1242, 375, align_front=af)
    assert "\n".join(o.to_kitti_format() for o in new_labels) == str(g[tag + "_labels"])
    assert labels[0].l == 3.9                       # inputs untouched (deep copies)


def test_scale_map_values():
    norm = load("stat_norm.norm")
    o3 = load("stat_norm.object_3d")
    obj = o3.Object3d("Car 0.00 0 1.20 600.00 150.00 700.00 220.00 1.50 1.60 3.90 2.00 1.60 12.00 0.30")
    m = norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)(obj, 1)
    assert m.shape == (1, 3)
    np.testing.assert_allclose(m.reshape(-1) * np.array([3.9, 1.5, 1.6]),
                               [3.9 + (5.149705924 - 4.401913719), 1.5 + (1.750965298 - 1.489997181),
                                1.6 + (1.934130886 - 1.788153724)], rtol=1e-15)
    assert np.array_equal(norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)(obj, 0), np.ones((1, 3)))


def test_convert_writes_a_rescaled_dataset(scene, tmp_path):
    g, norm, calib, labels, mapping, d = scene
    from PIL import Image
    src = tmp_path / "kitti"
    for sub in ("velodyne", "calib", "label_2", "image_2"):
        (src / "training" / sub).mkdir(parents=True)
    for split in ("train", "val", "trainval"):
        (src / (split + ".txt")).write_text("000000\n")
    g["velo"].tofile(str(src / "training" / "velodyne" / "000000.bin"))
    (src / "training" / "calib" / "000000.txt").write_text(str(g["calib"]))
    (src / "training" / "label_2" / "000000.txt").write_text(str(g["labels"]) + "\nDontCare -1 -1 -10 0 0 1 1 -1 -1 -1 -1000 -1000 -1000 -10")
    Image.new("RGB", (1242, 375)).save(str(src / "training" / "image_2" / "000000.png"))
    out = tmp_path / "out"
    norm.convert("kitti", "nusc", spath=str(src), dpath=str(out), use_car_sales_stats=True)
    root = out / "kitti_scaledto_nusc" / "training"
    raw = np.fromfile(str(root / "velodyne" / "000000.bin"), np.uint8)
    assert np.array_equal(sha(raw), g["ac0_af0_bin_sha"])
    assert (root / "label_2" / "000000.txt").read_text() == str(g["ac0_af0_labels"])
    assert os.path.islink(str(root / "calib")) and os.path.islink(str(root / "image_2"))
