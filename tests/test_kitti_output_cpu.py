"""kitti_output.kitti_lines (vectorised) == the literal per-box loop of save_kitti_format
(pointrcnn/tools/eval_rcnn.py:76-101), text for text, on random boxes."""
import os

import numpy as np

from conftest import load


def _literal(cfg, ku, calib, bbox3d, scores, img_shape):
    corners3d = ku.boxes3d_to_corners3d(bbox3d)
    img_boxes, _ = calib.corners3d_to_img_boxes(corners3d)
    img_boxes[:, 0] = np.clip(img_boxes[:, 0], 0, img_shape[1] - 1)
    img_boxes[:, 1] = np.clip(img_boxes[:, 1], 0, img_shape[0] - 1)
    img_boxes[:, 2] = np.clip(img_boxes[:, 2], 0, img_shape[1] - 1)
    img_boxes[:, 3] = np.clip(img_boxes[:, 3], 0, img_shape[0] - 1)
    img_boxes_w = img_boxes[:, 2] - img_boxes[:, 0]
    img_boxes_h = img_boxes[:, 3] - img_boxes[:, 1]
    box_valid_mask = np.logical_and(img_boxes_w < img_shape[1] * 0.8, img_boxes_h < img_shape[0] * 0.8)
    out = []
    for k in range(bbox3d.shape[0]):
        if box_valid_mask[k] == 0:
            continue
        x, z, ry = bbox3d[k, 0], bbox3d[k, 2], bbox3d[k, 6]
        beta = np.arctan2(z, x)
        alpha = -np.sign(beta) * np.pi / 2 + beta + ry
        out.append('%s -1 -1 %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f' %
                   (cfg.CLASSES, alpha, img_boxes[k, 0], img_boxes[k, 1], img_boxes[k, 2], img_boxes[k, 3],
                    bbox3d[k, 3], bbox3d[k, 4], bbox3d[k, 5], bbox3d[k, 0], bbox3d[k, 1], bbox3d[k, 2],
                    bbox3d[k, 6], scores[k]))
    return out


def test_vectorised_writer_equals_reference_loop(tmp_path):
    cfgm = load("config")
    cfgm.use_default_yaml("rcnn")
    ko, ku, cal, sk = load("kitti_output"), load("kitti_utils"), load("calibration"), load("synthetic_kitti")
    f = os.path.join(str(tmp_path), "calib.txt")
    open(f, "w").write("\n".join(sk.CALIB_LINES) + "\n")
    calib = cal.Calibration(f)
    rng = np.random.RandomState(0)
    total = 0
    for trial in range(60):
        n = rng.randint(0, 60)
        bbox3d = np.concatenate([rng.uniform(-30, 30, (n, 1)), rng.uniform(-1, 3, (n, 1)), rng.uniform(-5, 70, (n, 1)),
                                 rng.uniform(1, 2, (n, 3)), rng.uniform(-4, 4, (n, 1))], 1).astype(np.float32)
        scores = rng.randn(n).astype(np.float32)
        got = ko.kitti_lines(calib, bbox3d, scores, (375, 1242, 3))
        want = _literal(cfgm.cfg, ku, calib, bbox3d, scores, (375, 1242, 3)) if n else []
        assert got == want, trial
        total += len(got)
    assert total > 500
    ko.save_kitti_format(7, calib, bbox3d, str(tmp_path), scores, (375, 1242, 3))
    assert open(os.path.join(str(tmp_path), "000007.txt")).read().splitlines() == got
    assert ko.dump_empty_files(str(tmp_path), ["000007", "000008"]) == 1 and os.path.getsize(os.path.join(str(tmp_path), "000008.txt")) == 0
