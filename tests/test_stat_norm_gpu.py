"""GPU Statistical Normalization (csrc/stat_norm.cu, stat_norm/gpu_rescale.py) against the numpy path
stat_norm/norm.py -- itself pinned bit-exactly to the reference module by tests/test_stat_norm.py: the float32 rows of
the rescaled .bin and the ratios of every scene, and whole converted dataset trees, byte for byte."""
import os

import numpy as np
import pytest

from conftest import load
from test_stat_norm import GOLD

pytestmark = pytest.mark.gpu


def _objects(o3, rng, n_cars, extra=()):
    labels = []
    for _ in range(n_cars):
        x, z = rng.uniform(-15, 15), rng.uniform(6, 45)
        labels.append(o3.Object3d("%s 0.00 0 %.2f 600.00 150.00 700.00 220.00 %.2f %.2f %.2f %.2f %.2f %.2f %.2f" % (
            rng.choice(["Car", "Van"]), rng.uniform(-3.1, 3.1), rng.uniform(1.4, 2.0), rng.uniform(1.5, 1.9), rng.uniform(3.5, 5.0), x,
            rng.uniform(1.4, 1.8), z, rng.uniform(-3.1, 3.1))))
    for line in extra:
        labels.append(o3.Object3d(line))
    return labels


def _scene(calib, labels, rng, n):
    velo = np.stack([rng.uniform(0, 70, n), rng.uniform(-40, 40, n), rng.uniform(-3, 1, n), rng.uniform(0, 1, n)], 1)
    k = 0
    for obj in labels:
        m = int(rng.randint(0, 300))                      # some boxes stay empty (ratio 0)
        if m == 0 or k + m > n:
            continue
        loc = np.stack([rng.uniform(-obj.l / 2, obj.l / 2, m), rng.uniform(-obj.h, 0, m), rng.uniform(-obj.w / 2, obj.w / 2, m)], 1)
        c, s = np.cos(obj.ry), np.sin(obj.ry)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        velo[k:k + m, :3] = calib.project_rect_to_velo(loc @ R.T + obj.t)
        k += m
    return velo.astype(np.float32)


def _calib(tmp_path):
    ku = load("stat_norm.kitti_util")
    g = np.load(GOLD)
    p = tmp_path / "calib.txt"
    p.write_text(str(g["calib"]))
    return ku.Calibration(str(p)), str(g["calib"])


OPTIONS = [(False, False), (True, False), (False, True), (True, True)]     # (avoid_conflict, align_front), norm.py:205-240


@pytest.mark.parametrize("avoid_conflict,align_front", OPTIONS)
def test_rescale_scenes_gpu_equals_numpy_bytes(cuda, tmp_path, avoid_conflict, align_front):
    norm, gr, o3 = load("stat_norm.norm"), load("stat_norm.gpu_rescale"), load("stat_norm.object_3d")
    calib, _ = _calib(tmp_path)
    mapping = norm.get_scale_map(norm.germany_car_stats, norm.us_car_stats)
    rng = np.random.RandomState(5)
    scenes = []
    # overlapping boxes (a point in two boxes is emitted twice), a pedestrian (not rescaled), empty boxes, no boxes at all
    twin = ["Car 0.00 0 0.00 1 1 2 2 1.60 1.70 4.20 3.00 1.60 14.00 0.40", "Van 0.00 0 0.00 1 1 2 2 1.90 1.80 4.60 3.60 1.65 14.40 0.55",
            "Pedestrian 0.00 0 0.40 800.00 150.00 830.00 230.00 1.75 0.60 0.80 6.00 1.60 9.00 0.10"]
    for n, cars, extra in ((16384, 4, ()), (50000, 9, twin), (3000, 0, ()), (120000, 6, twin[:2]), (1, 1, ())):
        labels = _objects(o3, rng, cars, extra)
        scenes.append((_scene(calib, labels, rng, n), labels, calib))
    got = gr.rescale_scenes_gpu(mapping, scenes, cuda, avoid_conflict=avoid_conflict, align_front=align_front)
    assert len(got) == len(scenes)
    n_dup = 0
    seen_ratios = set()
    for (velo, labels, calib_), (rows, ratios) in zip(scenes, got):
        want_pts, want_ratios = norm.rescale_ptc(mapping, velo, labels, calib_, avoid_conflict=avoid_conflict, align_front=align_front)
        seen_ratios.update(float(r) for r in want_ratios)
        want = np.concatenate([want_pts, np.ones((want_pts.shape[0], 1), dtype=np.float32)], axis=1).astype(np.float32)   # norm.py:43
        assert ratios == want_ratios
        assert rows.shape == want.shape
        assert np.array_equal(rows.view(np.uint32), want.view(np.uint32))
        n_dup += rows.shape[0] - velo.shape[0]
    assert n_dup > 0                                          # the twin boxes really overlapped
    if avoid_conflict:
        assert len(seen_ratios - {0.0, 1.0}) > 0, seen_ratios    # the search really backed off for some box


@pytest.mark.parametrize("avoid_conflict,align_front", OPTIONS)
def test_convert_gpu_writes_the_same_dataset_as_convert(cuda, tmp_path, avoid_conflict, align_front):
    from PIL import Image
    norm, gr, o3 = load("stat_norm.norm"), load("stat_norm.gpu_rescale"), load("stat_norm.object_3d")
    calib, calib_txt = _calib(tmp_path)
    rng = np.random.RandomState(9)
    src = tmp_path / "kitti"
    for sub in ("velodyne", "calib", "label_2", "image_2"):
        (src / "training" / sub).mkdir(parents=True)
    names = ["%06d" % i for i in range(5)]
    for split in ("train", "val", "trainval"):
        (src / (split + ".txt")).write_text("\n".join(names) + "\n")
    for name in names:
        labels = _objects(o3, rng, int(rng.randint(0, 6)))
        _scene(calib, labels, rng, int(rng.randint(2000, 30000))).tofile(str(src / "training" / "velodyne" / (name + ".bin")))
        (src / "training" / "calib" / (name + ".txt")).write_text(calib_txt)
        (src / "training" / "label_2" / (name + ".txt")).write_text(
            "\n".join(o.to_kitti_format() for o in labels) + ("\n" if labels else "") +
            "DontCare -1 -1 -10 0 0 1 1 -1 -1 -1 -1000 -1000 -1000 -10")
        Image.new("RGB", (1242, 375)).save(str(src / "training" / "image_2" / (name + ".png")))
    norm.convert("kitti", "nusc", spath=str(src), dpath=str(tmp_path / "cpu"), use_car_sales_stats=True,
                 avoid_conflict=avoid_conflict, align_front=align_front)
    gr.convert_gpu("kitti", "nusc", spath=str(src), dpath=str(tmp_path / "gpu"), use_car_sales_stats=True, batch_size=2,
                   device=cuda, avoid_conflict=avoid_conflict, align_front=align_front)
    for sub in ("velodyne", "label_2"):
        a = tmp_path / "cpu" / "kitti_scaledto_nusc" / "training" / sub
        b = tmp_path / "gpu" / "kitti_scaledto_nusc" / "training" / sub
        assert sorted(os.listdir(str(a))) == sorted(os.listdir(str(b))) and len(os.listdir(str(a))) == 5
        for f in os.listdir(str(a)):
            assert open(str(a / f), "rb").read() == open(str(b / f), "rb").read(), (sub, f)
