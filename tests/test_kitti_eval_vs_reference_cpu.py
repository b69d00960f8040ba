"""Differential test of the evaluator mirror against the REFERENCE module itself (evaluate/eval2.py, numba CPU code,
imported from /root/reference with `rotate_iou` stubbed by the CPU oracle as tools/make_kitti_eval_fixture.py does).
Only where the reference tree exists (the build container); on the GPU box the committed golden vectors
(tests/test_kitti_eval_cpu.py, tests/test_kitti_eval_gpu.py) carry the parity.  Random small cases reach the corners a
single golden data set does not: tied scores and overlaps, ignored / DontCare combinations, empty sides, every mode
of compute_statistics_jit."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import load, ROOT

REF = "/root/reference/evaluate"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "eval2.py")), reason="reference tree not present")


@pytest.fixture(scope="module")
def both():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_kitti_eval_fixture as fx
    saved = sys.modules.get("rotate_iou")
    stub = types.ModuleType("rotate_iou")
    stub.rotate_iou_gpu_eval = fx.oracle_riou
    sys.modules["rotate_iou"] = stub
    sys.path.insert(0, REF)
    try:
        import eval2 as ref
    finally:
        sys.path.remove(REF)
        if saved is None:
            sys.modules.pop("rotate_iou", None)
        else:
            sys.modules["rotate_iou"] = saved
    ev = load("evaluate.eval2")
    old = ev.rotate_iou_gpu_eval
    ev.rotate_iou_gpu_eval = fx.oracle_riou
    yield ref, ev, fx
    ev.rotate_iou_gpu_eval = old
    sys.modules.pop("eval2", None)


def _boxes(rs, n):
    b = np.round(rs.uniform(0, 10, (n, 4)), 0 if rs.randint(2) else 3)
    b[:, 2:] = b[:, :2] + np.round(rs.uniform(0, 5, (n, 2)), 0 if rs.randint(2) else 3)      # degenerate sides included
    return b


def test_compute_statistics_every_mode(both):
    ref, ev, _ = both
    rs = np.random.RandomState(0)
    for it in range(400):
        ng, nd, ndc = int(rs.randint(0, 9)), int(rs.randint(0, 9)), int(rs.randint(0, 3))
        ov = rs.uniform(0, 1, (nd, ng))
        ov[rs.uniform(size=ov.shape) < 0.4] = 0.0
        if rs.randint(2):
            ov = np.round(ov, 1)                                    # tied overlaps
        gt = np.concatenate([_boxes(rs, ng), rs.uniform(-3, 3, (ng, 1))], 1)
        sc = np.round(rs.uniform(0, 1, (nd, 1)), 1 if rs.randint(2) else 6)     # tied scores
        dt = np.concatenate([_boxes(rs, nd), rs.uniform(-3, 3, (nd, 1)), sc], 1)
        ig, idt = rs.randint(-1, 2, ng).astype(np.int64), rs.randint(-1, 2, nd).astype(np.int64)
        dc = _boxes(rs, ndc)
        metric, mo = int(rs.randint(3)), float(rs.choice([0.25, 0.5, 0.7]))
        th = float(rs.choice([0.0, 0.3, 0.5, -100.0]))
        for cfp in (False, True):
            for aos in (False, True):
                a = ref.compute_statistics_jit(ov, gt, dt, ig, idt, dc, metric, mo, thresh=th, compute_fp=cfp, compute_aos=aos)
                b = ev.compute_statistics_jit(ov, gt, dt, ig, idt, dc, metric, mo, th, cfp, aos)
                assert a[:3] == b[:3], (it, cfp, aos)
                assert a[3] == b[3] or abs(a[3] - b[3]) < 1e-12, (it, cfp, aos, a[3], b[3])
                assert np.array_equal(np.asarray(a[4]), b[4]), (it, cfp, aos)


def test_image_box_overlap_and_thresholds(both):
    ref, ev, _ = both
    rs = np.random.RandomState(5)
    for _ in range(150):
        a, b = _boxes(rs, int(rs.randint(1, 6))), _boxes(rs, int(rs.randint(1, 6)))
        for crit in (-1, 0, 1, 2):
            assert np.array_equal(ref.image_box_overlap(a, b, crit), ev.image_box_overlap(a, b, crit), equal_nan=True)
    for _ in range(500):
        n = int(rs.randint(0, 60))
        sc = np.round(rs.uniform(0, 1, n), 2 if rs.randint(2) else 8)
        num_gt, ns = int(rs.randint(max(n, 1), n + 20)), int(rs.choice([11, 41]))
        assert np.array_equal(np.asarray(ref.get_thresholds(sc.copy(), num_gt, ns)),
                              np.asarray(ev.get_thresholds(sc.copy(), num_gt, ns)))


def test_clean_data_and_official_result_other_seeds(both, capsys):
    ref, ev, fx = both
    gts, dts = fx.make_annos(57, seed=7)                  # the reference needs more images than its 50 parts
    for g, d in zip(gts[:20], dts[:20]):
        for diff in (0, 1, 2):
            for cls in (0, 1, 2):
                r, m = ref.clean_data(g, d, cls, "kitti", diff), ev.clean_data(g, d, cls, "kitti", diff)
                assert r[0] == m[0] and list(r[1]) == list(m[1]) and list(r[2]) == list(m[2])
                assert np.array_equal(np.asarray(r[3], np.float64).reshape(-1, 4), np.asarray(m[3], np.float64).reshape(-1, 4))
    for cls in (0, [0, 1, 2]):
        r_txt, r_ret = ref.get_official_eval_result(gts, dts, cls, "kitti")
        m_txt, m_ret = ev.get_official_eval_result(gts, dts, cls, "kitti")
        assert r_txt == m_txt
        for k in r_ret:
            if k != "result":
                assert np.array_equal(np.float64(r_ret[k]), np.float64(m_ret[k]), equal_nan=True), k


def test_label_readers_equal_reference(tmp_path):
    """evaluate/kitti_common.py readers: the reference module (its unused `skimage` import stubbed) and the mirror on
    label files with / without the score column, an empty file, extra non-result files in the folder."""
    saved = {k: sys.modules.get(k) for k in ("skimage", "skimage.io", "kitti_common")}
    sk = types.ModuleType("skimage")
    sk.io = types.ModuleType("skimage.io")
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, sk.io
    sys.path.insert(0, REF)
    try:
        sys.modules.pop("kitti_common", None)
        import kitti_common as rkc
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    kc = load("evaluate.kitti_common")
    rs = np.random.RandomState(2)
    names = ["Car", "Pedestrian", "Cyclist", "Van", "DontCare", "Person_sitting"]
    for d, with_score in (("gt", False), ("dt", True)):
        os.makedirs(tmp_path / d)
        for i in (0, 3, 4, 11, 250):
            n = 0 if i == 4 else int(rs.randint(1, 9))
            lines = []
            for _ in range(n):
                v = rs.uniform(-50, 80, 12)
                line = "%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f" % (
                    names[rs.randint(len(names))], rs.uniform(0, 1), rs.randint(-1, 4), *v)
                lines.append(line + (" %.4f" % rs.uniform(-3, 3) if with_score else ""))
            (tmp_path / d / ("%06d.txt" % i)).write_text("\n".join(lines) + ("\n" if n and rs.randint(2) else ""))
        (tmp_path / d / "notes.txt").write_text("not a result file\n")
        (tmp_path / d / "1234567.txt").write_text("seven digits\n")
        for ids in (None, [3, 11], [4]):
            a, b = rkc.get_label_annos(str(tmp_path / d), ids), kc.get_label_annos(str(tmp_path / d), ids)
            assert len(a) == len(b) == (5 if ids is None else len(ids))
            for x, y in zip(a, b):
                assert set(x) == set(y)
                for k in x:
                    assert x[k].shape == y[k].shape and np.array_equal(x[k], y[k]), (d, ids, k)
                    assert x[k].dtype == y[k].dtype or x[k].size == 0, (d, ids, k, x[k].dtype, y[k].dtype)
    assert rkc.get_image_index_str(7) == kc.get_image_index_str(7) == "000007"


def test_evaluate_from_directories_equals_reference_evaluate_py(both, tmp_path, capsys):
    """evaluate/evaluate.py::evaluate (label / result directories + split file -> AP text and dict), the reference's
    driver imported live, against tools/kitti_ap.py::evaluate_dirs on the same files."""
    ref, ev, fx = both
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import kitti_ap
    gts, dts = fx.make_annos(57, seed=11)
    for d, annos, with_score in (("label_2", gts, False), ("data", dts, True)):
        os.makedirs(tmp_path / d)
        for i, a in enumerate(annos):
            lines = []
            for k in range(len(a["name"])):
                l, h, w = a["dimensions"][k]
                line = "%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f" % (
                    a["name"][k], a["truncated"][k], a["occluded"][k], a["alpha"][k], *a["bbox"][k], h, w, l,
                    *a["location"][k], a["rotation_y"][k])
                lines.append(line + (" %.4f" % a["score"][k] if with_score else ""))
            (tmp_path / d / ("%06d.txt" % i)).write_text("\n".join(lines))
    (tmp_path / "val.txt").write_text("\n".join(str(i) for i in range(57)) + "\n")
    # the reference driver: its kitti_common (skimage stubbed) and eval2 (rotate_iou stubbed with the CPU oracle)
    saved = {k: sys.modules.get(k) for k in ("skimage", "skimage.io", "kitti_common", "evaluate", "eval2", "rotate_iou")}
    sk = types.ModuleType("skimage")
    sk.io = types.ModuleType("skimage.io")
    stub = types.ModuleType("rotate_iou")
    stub.rotate_iou_gpu_eval = fx.oracle_riou
    sys.modules.update({"skimage": sk, "skimage.io": sk.io, "rotate_iou": stub, "eval2": ref})
    sys.modules.pop("kitti_common", None)
    sys.modules.pop("evaluate", None)
    sys.path.insert(0, REF)
    try:
        import evaluate as ref_driver
        want_txt, want = ref_driver.evaluate(str(tmp_path / "data"), label_split_file=str(tmp_path / "val.txt"),
                                             label_path=str(tmp_path / "label_2"), dataset="kitti", current_class=0)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    got_txt, got, _, _ = kitti_ap.evaluate_dirs(str(tmp_path / "label_2"), str(tmp_path / "data"), list(range(57)), "kitti", 0)
    assert got_txt == want_txt
    for k in want:
        if k != "result":
            assert np.array_equal(np.float64(want[k]), np.float64(got[k]), equal_nan=True), k


def test_other_datasets_difficulty_rules(both):
    """clean_data's per-dataset rules (kitti / argo / nusc / lyft / waymo) for all six difficulties and the three classes
    it knows (a class index beyond them is an IndexError on both sides), and the official result for two non-KITTI
    data sets."""
    ref, ev, fx = both
    gts, dts = fx.make_annos(57, seed=21)
    for dataset in ("kitti", "argo", "nusc", "lyft", "waymo"):
        for i in range(0, 57, 4):
            for diff in range(6):
                for cls in range(3):
                    r, m = ref.clean_data(gts[i], dts[i], cls, dataset, diff), ev.clean_data(gts[i], dts[i], cls, dataset, diff)
                    assert r[0] == m[0] and list(r[1]) == list(m[1]) and list(r[2]) == list(m[2]), (dataset, i, diff, cls)
    for mod in (ref, ev):
        with pytest.raises(IndexError):
            mod.clean_data(gts[0], dts[0], 3, "kitti", 0)
    for dataset in ("argo", "waymo"):
        r_txt, r_ret = ref.get_official_eval_result(gts, dts, 0, dataset)
        m_txt, m_ret = ev.get_official_eval_result(gts, dts, 0, dataset)
        assert r_txt == m_txt, dataset
        for k in r_ret:
            if k != "result":
                assert np.array_equal(np.float64(r_ret[k]), np.float64(m_ret[k]), equal_nan=True), (dataset, k)
