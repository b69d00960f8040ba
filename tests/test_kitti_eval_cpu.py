"""KITTI AP evaluator mirror (evaluate/eval2.py, csrc/kitti_eval.cu host functions) against golden vectors produced
by the REFERENCE evaluate/eval2.py (tools/make_kitti_eval_fixture.py): result text, official AP numbers, the full
precision / recall / orientation arrays of eval_class.  The rotated IoU is the CPU oracle on both sides here (no
GPU); tests/test_kitti_eval_gpu.py runs the same check with the sm_100a kernels."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")


def _fixture():
    fx = importlib.import_module("make_kitti_eval_fixture")
    z = np.load(os.path.join(GOLD, "kitti_eval.npz"))
    return fx, z, fx.unpack(z, "gt"), fx.unpack(z, "dt")


def check_against_golden(ev, z, gts, dts, exact=True):
    result, ret = ev.get_official_eval_result(gts, dts, 0, "kitti")
    close = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, rtol=0, atol=1e-9))
    for k in ("Car_3d_easy", "Car_3d_moderate", "Car_3d_hard", "Car_bev_easy", "Car_bev_moderate", "Car_bev_hard",
              "Car_image_easy", "Car_image_moderate", "Car_image_hard"):
        assert close(np.float64(ret[k]), z["ret_" + k]), (k, ret[k], z["ret_" + k])
    for key, val in ret["result"][0].items():
        tag = "res%d_" % (0 if "0.70, 0.70, 0.70" in key else 1)
        for name, arr in val.items():
            assert close(np.asarray(arr, np.float64), z[tag + name]), (key, name)
    assert result == open(os.path.join(GOLD, "kitti_eval.txt")).read()
    mo = np.stack([np.array([[0.7, 0.5, 0.5, 0.7, 0.5]] * 3),
                   np.array([[0.7, 0.5, 0.5, 0.7, 0.5], [0.5, 0.25, 0.25, 0.5, 0.25], [0.5, 0.25, 0.25, 0.5, 0.25]])], 0)[:, :, [0]]
    for name, metric, aos in (("cls3d", 2, False), ("clsbb", 0, True)):
        got = ev.eval_class(gts, dts, [0], "kitti", [0, 1, 2, 3, 4, 5], metric, mo, compute_aos=aos)
        for k in ("recall", "precision", "orientation"):
            a, b = got[k], z[name + "_" + k]
            assert a.shape == b.shape
            assert np.array_equal(np.isnan(a), np.isnan(b)), (name, k)
            m = ~np.isnan(a)
            if k == "orientation" or not exact:
                assert np.allclose(a[m], b[m], rtol=0, atol=1e-12), (name, k)       # cos() summation order
            else:
                assert np.array_equal(a[m], b[m]), (name, k)


def test_evaluator_equals_reference_golden(monkeypatch):
    fx, z, gts, dts = _fixture()
    ev = load("evaluate.eval2")
    monkeypatch.setattr(ev, "rotate_iou_gpu_eval", fx.oracle_riou)
    check_against_golden(ev, z, gts, dts, exact=True)


def test_single_image_functions_and_readers(tmp_path, monkeypatch):
    fx, z, gts, dts = _fixture()
    ev, kc = load("evaluate.eval2"), load("evaluate.kitti_common")
    monkeypatch.setattr(ev, "rotate_iou_gpu_eval", fx.oracle_riou)
    # image_box_overlap against a direct numpy evaluation, all criteria, empty sides
    rng = np.random.RandomState(1)
    a = np.sort(rng.uniform(0, 100, (7, 2, 2)), axis=1).transpose(0, 2, 1).reshape(7, 4)[:, [0, 2, 1, 3]]
    b = np.sort(rng.uniform(0, 100, (5, 2, 2)), axis=1).transpose(0, 2, 1).reshape(5, 4)[:, [0, 2, 1, 3]]
    for crit in (-1, 0, 1, 2):
        got = ev.image_box_overlap(a, b, crit)
        iw = np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0])
        ih = np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1])
        inter = np.where((iw > 0) & (ih > 0), iw * ih, 0.0)
        aa = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
        ab = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
        ua = {-1: aa + ab - inter, 0: aa + 0 * ab, 1: ab + 0 * aa, 2: np.ones_like(inter)}[crit]
        assert np.allclose(got, np.where(inter > 0, inter / ua, 0.0), rtol=1e-15, atol=0)
    assert ev.image_box_overlap(np.zeros((0, 4)), b).shape == (0, 5)
    # per-image API == the part-wise driver on a one-image part
    i = int(np.argmax([len(g["name"]) * len(d["name"]) for g, d in zip(gts, dts)]))
    ov, _, _, _ = ev.calculate_iou_partly([dts[i]], [gts[i]], 2, 1)
    gd, dd, ig, idt, dc, dcn, nvalid = ev._prepare_data([gts[i]], [dts[i]], 0, "kitti", 1)
    tp, _, _, _, th = ev.compute_statistics_jit(ov[0], gd[0], dd[0], ig[0], idt[0], dc[0], 2, 0.5, 0.0, False)
    assert tp == len(th) and tp >= 1 and set(np.round(th, 9)) <= set(np.round(dts[i]["score"], 9))
    tp2, fp2, fn2, sim, _ = ev.compute_statistics_jit(ov[0], gd[0], dd[0], ig[0], idt[0], dc[0], 2, 0.5, -100.0, True, True)
    assert tp2 == tp and tp2 + fn2 == nvalid and fp2 >= 0 and 0 <= sim <= tp2 + 1e-12
    # label readers: write -> read round trip in the KITTI text format (h w l on disk, l h w in memory)
    g = gts[i]
    path = tmp_path / "000003.txt"
    with open(path, "w") as f:
        for k in range(len(g["name"])):
            l, h, w = g["dimensions"][k]
            f.write("%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n" % (
                g["name"][k], g["truncated"][k], g["occluded"][k], g["alpha"][k], *g["bbox"][k], h, w, l, *g["location"][k],
                g["rotation_y"][k]))
    an = kc.get_label_annos(str(tmp_path))[0]
    assert list(an["name"]) == list(g["name"]) and np.allclose(an["dimensions"], g["dimensions"], atol=0.006)
    assert an["score"].shape == (len(g["name"]),) and np.all(an["score"] == 0)
    assert kc.get_label_annos(str(tmp_path), [3])[0]["bbox"].shape == (len(g["name"]), 4)
