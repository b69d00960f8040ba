"""KITTI result files: mirror of save_kitti_format (pointrcnn/tools/eval_rcnn.py:76-101) and of the "dump empty
files" loop (:638-649), for drivers that do not go through eval_rcnn.py (tools/eval_fast.py)."""
import os

import numpy as np

from . import kitti_utils
from .config import cfg


def kitti_lines(calib, bbox3d, scores, img_shape, cls_name=None):
    """-> list of the text lines eval_rcnn.py writes for one scene (same arithmetic, same '%.4f' formatting)."""
    cls_name = cfg.CLASSES if cls_name is None else cls_name
    if bbox3d.shape[0] == 0:
        return []
    corners3d = kitti_utils.boxes3d_to_corners3d(bbox3d)
    img_boxes, _ = calib.corners3d_to_img_boxes(corners3d)
    img_boxes[:, 0] = np.clip(img_boxes[:, 0], 0, img_shape[1] - 1)
    img_boxes[:, 1] = np.clip(img_boxes[:, 1], 0, img_shape[0] - 1)
    img_boxes[:, 2] = np.clip(img_boxes[:, 2], 0, img_shape[1] - 1)
    img_boxes[:, 3] = np.clip(img_boxes[:, 3], 0, img_shape[0] - 1)
    img_boxes_w = img_boxes[:, 2] - img_boxes[:, 0]
    img_boxes_h = img_boxes[:, 3] - img_boxes[:, 1]
    box_valid_mask = np.logical_and(img_boxes_w < img_shape[1] * 0.8, img_boxes_h < img_shape[0] * 0.8)
    # vectorised over the boxes of the scene; the float32 operation order of eval_rcnn.py:93-95 is kept element-wise:
    #   beta = arctan2(z, x) ; alpha = -sign(beta) * pi / 2 + beta + ry
    b32 = np.asarray(bbox3d, np.float32)
    beta = np.arctan2(b32[:, 2], b32[:, 0])
    alpha = -np.sign(beta) * np.pi / 2 + beta + b32[:, 6]
    rows = np.concatenate((alpha.reshape(-1, 1), img_boxes[:, 0:4], b32[:, 3:6], b32[:, 0:3], b32[:, 6:7],
                           np.asarray(scores).reshape(-1, 1)), axis=1).astype(np.float64)     # '%.4f' formats the double
    fmt = cls_name + ' -1 -1' + ' %.4f' * 13
    # tolist(): Python floats of the same doubles -- formatting numpy scalars one by one costs 2.5x as much
    return [fmt % tuple(r) for r in rows[box_valid_mask].tolist()]


def save_kitti_format(sample_id, calib, bbox3d, kitti_output_dir, scores, img_shape):
    """eval_rcnn.py:76-101.  Like the reference's call site (:614-629) a scene without a detection above the score
    threshold gets no file here; dump_empty_files() adds the empty ones at the end."""
    lines = kitti_lines(calib, bbox3d, scores, img_shape)
    with open(os.path.join(kitti_output_dir, '%06d.txt' % sample_id), 'w') as f:
        if lines:
            f.write('\n'.join(lines) + '\n')


def dump_empty_files(final_output_dir, image_idx_list):
    """eval_rcnn.py:638-649: an empty result file for every id of the split without one."""
    n = 0
    for idx in image_idx_list:
        cur_file = os.path.join(final_output_dir, '%s.txt' % idx)
        if not os.path.exists(cur_file):
            open(cur_file, 'w').close()
            n += 1
    return n
