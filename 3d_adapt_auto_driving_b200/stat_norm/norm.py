"""Mirror of stat_norm/norm.py -- Statistical Normalization of a KITTI-format scene.

Same callables and signatures as the reference:
    single_scale, get_scale_map, format_lidar_data, save_labels, gen_obj_box_ptc, refine,
    postprocessing, regenerate_labels, scale_labels, rescale_ptc, convert, launch_rescale
The arithmetic is float64 numpy in the reference's operation order (stat_norm/norm.py:186-244 for
the point rescale, :154-183 for the labels), so rescaled coordinates are bit-identical to the
reference's on the same BLAS (tests/test_stat_norm.py checks against outputs of the reference
module itself, tools/make_statnorm_fixture.py).

Differences that the drop-in absorbs:
  * no import side effects (the reference's config_path.py prints and creates
    ~/scratch/driving_datasets at import, config_path.py:3-8): dataset roots are arguments;
  * `postprocessing` works on NumPy 2: the reference builds its occupancy map as
    np.ones(uint8) * -1 (norm.py:134), an OverflowError since NumPy 2.0; the map is int16 here,
    which is what NumPy 1.x value-based casting produced.
"""
import copy
import json
import multiprocessing as _mp
import os
import shutil
from itertools import chain

import numpy as np

from .kitti_util import Calibration, load_velo_scan
from .object_3d import read_label

car_sales_path = os.path.join(os.path.dirname(os.path.realpath(__file__)), "car_sales")


def load_json(fname):
    with open(fname, "r") as f:
        return json.load(f)


us_car_stats = load_json(os.path.join(car_sales_path, "us.json"))
germany_car_stats = load_json(os.path.join(car_sales_path, "germany.json"))
car_stats_external = {"kitti": germany_car_stats, "argo_new": us_car_stats, "nusc": us_car_stats,
                      "lyft": us_car_stats, "waymo": us_car_stats}
# config_path.py:24-40: the experiment datasets; "argo" has no car-sales table (norm.py:34-39 names it "argo_new"),
# so convert("argo", ..., use_car_sales_stats=True) raises KeyError exactly like the reference
datasets = ("kitti", "argo", "nusc", "lyft", "waymo")


def format_lidar_data(x, dst):
    """(N,3) velodyne coordinates -> KITTI .bin (N,4) float32 with intensity 1.0 (norm.py:42-45)."""
    x = np.concatenate([x, np.ones((x.shape[0], 1), dtype=np.float32)], axis=1).astype(np.float32)
    x.reshape(-1).tofile(dst)


def save_labels(labels, dst):
    with open(dst, "w") as f:
        f.write("\n".join(obj.to_kitti_format() for obj in labels))


def single_scale(x, src, dst, ratio=1):
    return x + (dst["mean"] - src["mean"]) * ratio


def get_scale_map(src, dst):
    """-> f(obj, ratio) = per-axis scale factors (1,3) in box-frame order (l, h, w) (norm.py:54-64)."""
    return lambda x, ratio: (np.array([
        single_scale(x.l, src["length"], dst["length"], ratio),
        single_scale(x.h, src["height"], dst["height"], ratio),
        single_scale(x.w, src["width"], dst["width"], ratio),
    ]) / np.array([x.l, x.h, x.w])).reshape(1, 3)


def roty(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def gen_obj_box_ptc(obj):
    """(8,3) corners of the label box in rect-camera coordinates (norm.py:94-115)."""
    l, w, h = obj.l, obj.w, obj.h
    x_corners = [l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2]
    y_corners = [-h, -h, -h, -h, 0, 0, 0, 0]
    z_corners = [w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2]
    corners_3d = np.dot(roty(obj.ry), np.vstack([x_corners, y_corners, z_corners]))
    corners_3d[0, :] = corners_3d[0, :] + obj.t[0]
    corners_3d[1, :] = corners_3d[1, :] + obj.t[1]
    corners_3d[2, :] = corners_3d[2, :] + obj.t[2]
    return np.transpose(corners_3d)


def refine(obj, calib, w, h):
    """2-D box = image-plane bounding box of the (rescaled) 3-D box, clipped to the image (norm.py:118-130)."""
    uv = calib.project_rect_to_image2(gen_obj_box_ptc(obj))
    bbox = list(chain(np.min(uv, axis=0).tolist()[0:2], np.max(uv, axis=0).tolist()[0:2]))
    obj.box2d = np.array([max(0, bbox[0]), max(0, bbox[1]), min(w, bbox[2]), min(h, bbox[3])])
    return obj


def postprocessing(objs, w, h):
    """re-derive the truncation field from how much of each 2-D box is covered by nearer boxes
    (painter's algorithm on a per-pixel owner map, far to near; norm.py:133-146)."""
    owner = np.ones((h, w), dtype=np.int16) * -1
    objs = sorted(objs, key=lambda x: x.t[2], reverse=True)
    for i, obj in enumerate(objs):
        owner[int(round(obj.box2d[1])):int(round(obj.box2d[3])), int(round(obj.box2d[0])):int(round(obj.box2d[2]))] = i
    unique, counts = np.unique(owner, return_counts=True)
    counts = dict(zip(unique, counts))
    for i, obj in enumerate(objs):
        if i not in counts.keys():
            counts[i] = 0
        occlusion = 1.0 - counts[i] / (obj.box2d[3] - obj.box2d[1]) / (obj.box2d[2] - obj.box2d[0])
        obj.trucation = int(np.clip(occlusion * 4, 0, 3))
    return objs


def regenerate_labels(objs, calib, w, h):
    for i in range(len(objs)):
        objs[i] = refine(objs[i], calib, w, h)
    return postprocessing(objs, w, h)


def _front_alignment_shifts(obj, l, w):
    """[(shift, angle), ...] that keep the face of the box nearest to the sensor where it was
    (norm.py:162-178 and :221-239: the same rule moves the label and the points)."""
    out = []
    dist = np.linalg.norm(obj.t)
    alpha = np.arctan2(np.sin(obj.alpha), np.cos(obj.alpha))
    if np.abs(np.sin(alpha)) * dist > obj.l / 2.0:
        shift = (obj.l - l) / 2.0
        angle = -obj.ry if 0 < alpha else -obj.ry + np.pi
        out.append((shift, angle))
    if np.abs(np.cos(alpha)) * dist > obj.w / 2.0:
        shift = (obj.w - w) / 2.0
        angle = -obj.ry - np.pi / 2.0 if -np.pi / 2.0 < alpha < np.pi / 2.0 else -obj.ry + np.pi / 2.0
        out.append((shift, angle))
    return out


def scale_labels(objs, mapping, ratios, calib, w0, h0, align_front=False, rescaled_classes=("Car", "Van")):
    new_obj = []
    cnt = 0
    for obj in objs:
        _obj = copy.deepcopy(obj)
        if obj.cls_type in rescaled_classes:
            l, h, w = (np.array([obj.l, obj.h, obj.w]) * mapping(obj, ratios[cnt]).reshape(-1)).tolist()
            if align_front:
                for shift, angle in _front_alignment_shifts(obj, l, w):
                    _obj.t[0] += shift * np.cos(angle)
                    _obj.t[2] += shift * np.sin(angle)
            _obj.l, _obj.h, _obj.w = l, h, w
            cnt += 1
        new_obj.append(_obj)
    return regenerate_labels(new_obj, calib, w0, h0)


def _box_frame(obj):
    """rotation whose columns are the box axes in rect-camera coordinates: (p - t) @ R = box-frame coordinates
    (x along the length, y negative upwards from the bottom face, z along the width)"""
    c, s = np.cos(obj.ry), np.sin(obj.ry)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _between(v, lo, hi):
    return (v > lo) & (v < hi)


def _largest_safe_ratio(mapping, obj, local, inside, ground_clear_count):
    """avoid_conflict (norm.py:205-216): take the full change (ratio 1) and back off in steps of 0.1 until the scaled
    patch's bounding volume, ignoring the lowest 0.5 m, holds fewer than 10 points more than the original box did"""
    for ratio in np.arange(1, -0.1, -0.1):
        scaled = local[inside] * mapping(obj, ratio)
        lo, hi = np.min(scaled, axis=0), np.max(scaled, axis=0)
        swallowed = _between(local[:, 0], lo[0], hi[0]) & _between(local[:, 1], lo[1], -0.5) & _between(local[:, 2], lo[2], hi[2])
        if np.sum(swallowed) - ground_clear_count < 10:
            break
    return ratio, scaled


def rescale_ptc(mapping, velo, labels, calib, avoid_conflict=False, align_front=False, rescaled_classes=("Car", "Van")):
    """velo (N,4) float32, labels [Object3d], calib -> ((N,3) float64 velodyne coordinates, ratios).
    Points strictly inside a Car / Van box are scaled about the box's bottom-face centre along its
    own axes; output order is [patch of box 0, patch of box 1, ..., untouched points] (norm.py:186-244)."""
    ptc = calib.project_velo_to_rect(velo[:, :3])
    untouched = np.ones(ptc.shape[0]).astype(bool)
    patches, ratios = [], []
    for obj in (o for o in labels if o.cls_type in rescaled_classes):
        R = _box_frame(obj)
        local = np.dot(ptc - obj.t, R)
        in_footprint = _between(local[:, 0], -obj.l / 2.0, obj.l / 2.0), _between(local[:, 2], -obj.w / 2.0, obj.w / 2.0)
        above_floor = local[:, 1] > -obj.h
        inside = in_footprint[0] & above_floor & (local[:, 1] < 0) & in_footprint[1]
        ratio = 0
        if np.sum(inside) > 0:
            untouched[inside] = False
            if avoid_conflict:
                ground_clear = in_footprint[0] & above_floor & (local[:, 1] < -0.5) & in_footprint[1]
                ratio, scaled = _largest_safe_ratio(mapping, obj, local, inside, np.sum(ground_clear))
            else:
                ratio = 1
                scaled = local[inside] * mapping(obj, ratio)
            patch = np.dot(scaled, R.T) + obj.t
            if align_front:
                l, h, w = (np.array([obj.l, obj.h, obj.w]) * mapping(obj, ratio).reshape(-1)).tolist()
                for shift, angle in _front_alignment_shifts(obj, l, w):
                    patch[:, 0] += shift * np.cos(angle)
                    patch[:, 2] += shift * np.sin(angle)
            patches.append(patch)
        ratios.append(ratio)
    return calib.project_rect_to_velo(np.concatenate(patches + [ptc[untouched]], axis=0)), ratios


def get_image_size(path):
    from PIL import Image
    with open(os.path.join(path, "train.txt")) as f:
        sample_img_name = f.readlines()[0]
    return Image.open(os.path.join(path, "training", "image_2", "%s.png" % sample_img_name.rstrip())).size


def convert(src, dst, spath=None, dpath=None, image_folder="image_2", calib_folder="calib", label_folder="label_2",
            use_car_sales_stats=False, avoid_conflict=False, align_front=False, rescaled_classes=("Car", "Van"),
            dataset_paths=None):
    """rescale a whole KITTI-format dataset `src` towards the car-size statistics of `dst`
    (norm.py:247-307).  `dataset_paths` ({name: root}) replaces the reference's config_path module."""
    assert src in datasets and dst in datasets
    dataset_paths = dataset_paths or {}
    spath = spath or dataset_paths[src]
    if use_car_sales_stats:
        mapping = get_scale_map(car_stats_external[src], car_stats_external[dst])
    else:
        mapping = get_scale_map(load_json(os.path.join(dataset_paths[src], "label_stats_train.json")),
                                load_json(os.path.join(dataset_paths[dst], "label_stats_train.json")))
    w, h = get_image_size(spath)
    if dpath is None:
        raise ValueError("dpath (where the rescaled datasets are written) must be given")
    root = os.path.join(dpath, "%s_scaledto_%s" % (src, dst))
    os.makedirs(root, exist_ok=True)
    for split in ["train", "val", "trainval"]:
        shutil.copyfile(os.path.join(spath, "%s.txt" % split), os.path.join(root, "%s.txt" % split))
    root = os.path.join(root, "training")
    os.makedirs(root, exist_ok=True)
    for link, folder in (("image_2", image_folder), ("calib", calib_folder)):
        target = os.path.join(root, link)
        if os.path.islink(target) or os.path.exists(target):
            os.remove(target)
        os.symlink(os.path.join(spath, "training", folder), target)
    os.makedirs(os.path.join(root, "velodyne"), exist_ok=True)
    os.makedirs(os.path.join(root, label_folder), exist_ok=True)
    with open(os.path.join(spath, "trainval.txt")) as f:
        names = [x.strip() for x in f.readlines()]
    for name in names:
        ptc = load_velo_scan(os.path.join(spath, "training", "velodyne", "%s.bin" % name))
        calib = Calibration(os.path.join(spath, "training", calib_folder, "%s.txt" % name))
        labels = [x for x in read_label(os.path.join(spath, "training", label_folder, "%s.txt" % name))
                  if x.cls_type != "DontCare"]
        new_ptc, ratios = rescale_ptc(mapping, ptc, labels, calib, avoid_conflict=avoid_conflict,
                                      align_front=align_front, rescaled_classes=rescaled_classes)
        format_lidar_data(new_ptc, os.path.join(root, "velodyne", "%s.bin" % name))
        labels = scale_labels(labels, mapping, ratios, calib, w, h, align_front=align_front,
                              rescaled_classes=rescaled_classes)
        save_labels(labels, os.path.join(root, label_folder, "%s.txt" % name))


def launch_rescale(dataset_paths, **kwargs):
    """one spawned process per ordered (src, dst) pair of the given datasets (norm.py:310-320)."""
    mp = _mp.get_context('spawn')
    processes = []
    for src in dataset_paths:
        for dst in dataset_paths:
            if src != dst:
                p = mp.Process(target=convert, args=(src, dst), kwargs=dict(kwargs, dataset_paths=dataset_paths))
                p.start()
                processes.append(p)
    for p in processes:
        p.join()
