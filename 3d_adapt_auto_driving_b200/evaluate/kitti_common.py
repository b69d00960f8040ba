"""Mirror of evaluate/kitti_common.py:307-360: KITTI label / result text files -> annotation dicts (the reference
module also imports skimage for image helpers that the evaluator never uses; they are not mirrored)."""
import pathlib
import re

import numpy as np


def get_image_index_str(img_idx):
    return "{:06d}".format(img_idx)


def get_label_anno(label_path):
    """kitti_common.py:307-343.  dimensions are converted from the file's h w l to l h w (camera)."""
    annotations = {}
    with open(label_path, 'r') as f:
        lines = f.readlines()
    content = [line.strip().split(' ') for line in lines]
    annotations['name'] = np.array([x[0] for x in content])
    annotations['truncated'] = np.array([float(x[1]) for x in content])
    annotations['occluded'] = np.array([int(x[2]) for x in content])
    annotations['alpha'] = np.array([float(x[3]) for x in content])
    annotations['bbox'] = np.array([[float(info) for info in x[4:8]] for x in content]).reshape(-1, 4)
    annotations['dimensions'] = np.array([[float(info) for info in x[8:11]] for x in content]).reshape(-1, 3)[:, [2, 0, 1]]
    annotations['location'] = np.array([[float(info) for info in x[11:14]] for x in content]).reshape(-1, 3)
    annotations['rotation_y'] = np.array([float(x[14]) for x in content]).reshape(-1)
    if len(content) != 0 and len(content[0]) == 16:  # have score
        annotations['score'] = np.array([float(x[15]) for x in content])
    else:
        annotations['score'] = np.zeros([len(annotations['bbox'])])
    return annotations


def get_label_annos(label_folder, image_ids=None):
    """kitti_common.py:345-360."""
    if image_ids is None:
        filepaths = pathlib.Path(label_folder).glob('*.txt')
        prog = re.compile(r'^\d{6}.txt$')
        filepaths = filter(lambda f: prog.match(f.name), filepaths)
        image_ids = sorted(int(p.stem) for p in filepaths)
    if not isinstance(image_ids, list):
        image_ids = list(range(image_ids))
    label_folder = pathlib.Path(label_folder)
    return [get_label_anno(label_folder / (get_image_index_str(idx) + '.txt')) for idx in image_ids]
