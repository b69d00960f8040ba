"""KITTI label / result text files -> the annotation dicts the evaluator consumes: the interface of
evaluate/kitti_common.py:307-360 (`get_label_anno`, `get_label_annos`, `get_image_index_str`).  The reference module
also imports skimage for image helpers the evaluator never uses; those are not provided.

One pass per file: the numeric columns of all rows are parsed into a single float64 matrix and the dict fields are
slices of it (the reference walks the token lists once per field).  Field names, dtypes, shapes and the h w l -> l h w
reordering of `dimensions` are the reference's; tests/test_kitti_eval_vs_reference_cpu.py compares the two readers on
files with and without a score column, empty files and directory listings."""
import os
import re

import numpy as np

_RESULT_NAME = re.compile(r'^\d{6}.txt$')
# column layout of a KITTI object line after the class name (token 0)
_TRUNC, _OCC, _ALPHA, _BBOX, _HWL, _XYZ, _RY, _SCORE = 0, 1, 2, slice(3, 7), slice(7, 10), slice(10, 13), 13, 14


def get_image_index_str(img_idx):
    return "%06d" % img_idx


def get_label_anno(label_path):
    with open(label_path, 'r') as f:
        rows = [line.strip().split(' ') for line in f.readlines()]
    n = len(rows)
    has_score = n > 0 and len(rows[0]) == 16             # decided by the first line, as the reference does
    width = 15 if has_score else 14
    num = np.array([r[1:1 + width] for r in rows], dtype=np.float64).reshape(n, width)
    return {
        'name': np.array([r[0] for r in rows]),
        'truncated': num[:, _TRUNC].copy(),
        'occluded': np.array([int(r[2]) for r in rows]),  # int(): a non-integer occlusion field raises, as there
        'alpha': num[:, _ALPHA].copy(),
        'bbox': num[:, _BBOX].copy(),
        'dimensions': num[:, _HWL][:, [2, 0, 1]],         # file order h w l -> l h w (camera convention)
        'location': num[:, _XYZ].copy(),
        'rotation_y': num[:, _RY].copy(),
        'score': num[:, _SCORE].copy() if has_score else np.zeros([n]),
    }


def get_label_annos(label_folder, image_ids=None):
    """image_ids: None = every ######.txt of the folder in ascending order; an int k = ids 0..k-1; or a list."""
    folder = os.fspath(label_folder)
    if image_ids is None:
        image_ids = sorted(int(name[:-4]) for name in os.listdir(folder) if _RESULT_NAME.match(name))
    elif not isinstance(image_ids, list):
        image_ids = list(range(image_ids))
    return [get_label_anno(os.path.join(folder, get_image_index_str(i) + '.txt')) for i in image_ids]
