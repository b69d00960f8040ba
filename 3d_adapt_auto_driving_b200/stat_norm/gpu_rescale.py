"""Statistical Normalization on the GPU for whole batches of scenes: the point rescale of
stat_norm/norm.py (rescale_ptc + format_lidar_data, norm.py:186-244, 42-45) behind pn2_stat_rescale_f64
(csrc/stat_norm.cu), and `convert_gpu`, the dataset driver of norm.py:247-307 on top of it.

All four combinations of convert's options: avoid_conflict (the per-box search for the largest ratio whose scaled
patch does not swallow foreign points, norm.py:205-216) runs on the device as one min / max reduction over the in-box
points plus one count per candidate ratio; align_front (norm.py:219-240) is two scalar shifts per box.  Everything numpy
builds from scalars -- cos / sin of ry, inv(R0), the scale factors mapping(obj, ratio) for the eleven candidate ratios
of np.arange(1, -0.1, -0.1), the front-alignment shifts shift * cos(angle) -- is computed here by the same numpy calls
and handed to the kernel, which reproduces the float64 np.dot chains bit for bit; the rows written to the .bin files
are byte-identical to norm.convert's (tests/test_stat_norm_gpu.py).  Label rescaling (scale_labels, a
few objects per scene) is the unchanged host code."""
import ctypes
import os
import shutil

import numpy as np
import torch

from .. import cabi
from ..cabi import i32, ptr
from . import norm
from .kitti_util import Calibration, load_velo_scan
from .object_3d import read_label


def _scene_mats(calib):
    m = np.empty((42,), np.float64)
    m[0:12] = np.transpose(calib.V2C).reshape(-1)
    m[12:21] = np.asarray(calib.R0, np.float64).reshape(-1)
    m[21:30] = np.linalg.inv(calib.R0).reshape(-1)              # kitti_util.py:151
    m[30:42] = np.transpose(calib.C2V).reshape(-1)
    return m


def _box_params(obj, mapping):
    c, s = np.cos(obj.ry), np.sin(obj.ry)                           # norm.py:195-196
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    p = np.empty((18,), np.float64)
    p[0:3] = np.asarray(obj.t, np.float64)
    p[3:12] = R.reshape(-1)
    p[12], p[13], p[14] = obj.l / 2.0, obj.h, obj.w / 2.0
    p[15:18] = mapping(obj, 1).reshape(-1)
    return p


RATIOS = np.arange(1, -0.1, -0.1)          # norm.py:206, the candidates of the avoid_conflict search


def _box_options(obj, mapping, avoid_conflict, align_front):
    """(78,) float64: mapping(obj, ratio) for the eleven candidate ratios, the front-alignment (dx, dz) pairs for each
    of them (the reference's own scalar expressions, norm.py:219-240 via norm._front_alignment_shifts) and their number."""
    o = np.zeros((78,), np.float64)
    n_pairs = 0
    for r, ratio in enumerate(RATIOS if avoid_conflict else RATIOS[:1]):
        ratio = ratio if avoid_conflict else 1
        factors = mapping(obj, ratio).reshape(-1)
        o[3 * r:3 * r + 3] = factors
        if align_front:
            l, h, w = (np.array([obj.l, obj.h, obj.w]) * factors).tolist()
            pairs = norm._front_alignment_shifts(obj, l, w)
            n_pairs = len(pairs)
            for k, (shift, angle) in enumerate(pairs):
                o[33 + 4 * r + 2 * k] = shift * np.cos(angle)
                o[33 + 4 * r + 2 * k + 1] = shift * np.sin(angle)
    o[77] = n_pairs
    return o


@torch.no_grad()
def rescale_scenes_gpu(mapping, scenes, device=None, rescaled_classes=("Car", "Van"), avoid_conflict=False, align_front=False):
    """scenes: list of (velo (N,4) float32, labels [Object3d], calib) -> list of (bin_rows (M,4) float32, ratios):
    exactly `format_lidar_data(rescale_ptc(mapping, velo, labels, calib, avoid_conflict, align_front)[0])`'s rows and
    rescale_ptc's ratios."""
    if not scenes:
        return []
    device = device or torch.device("cuda", torch.cuda.current_device())
    b = len(scenes)
    sizes = [int(v.shape[0]) for v, _, _ in scenes]
    offsets = np.zeros((b + 1,), np.int64)
    offsets[1:] = np.cumsum(sizes)
    raw = torch.empty((int(offsets[-1]), 4), dtype=torch.float32).pin_memory()
    mats = np.empty((b, 42), np.float64)
    boxes, box_offsets, rescaled, options = [], [0], [], []
    with_options = avoid_conflict or align_front
    for k, (velo, labels, calib) in enumerate(scenes):
        raw[int(offsets[k]):int(offsets[k + 1])] = torch.from_numpy(np.ascontiguousarray(velo, np.float32))
        mats[k] = _scene_mats(calib)
        objs = [o for o in labels if o.cls_type in rescaled_classes]
        rescaled.append(objs)
        boxes.extend(_box_params(o, mapping) for o in objs)
        if with_options:
            options.extend(_box_options(o, mapping, avoid_conflict, align_front) for o in objs)
        box_offsets.append(len(boxes))
    nb = len(boxes)
    boxes_np = np.stack(boxes) if nb else np.zeros((1, 18), np.float64)
    cap = (max(sizes) + 1023) // 1024 * 1024
    d_raw = raw.to(device, non_blocking=True)
    d_off = torch.from_numpy(offsets).to(device)
    d_mats = torch.from_numpy(mats).to(device)
    d_boxes = torch.from_numpy(boxes_np).to(device)
    d_boff = torch.from_numpy(np.array(box_offsets, np.int32)).to(device)
    rect = torch.empty((b, cap, 3), dtype=torch.float64, device=device)
    untouched = torch.empty((b, cap), dtype=torch.uint8, device=device)
    box_counts = torch.zeros((max(nb, 1),), dtype=torch.int32, device=device)
    box_ratio = torch.zeros((max(nb, 1),), dtype=torch.int32, device=device)
    d_opts = torch.from_numpy(np.stack(options) if options else np.zeros((1, 78), np.float64)).to(device) if with_options else None
    cap_out = cap + cap // 4                                     # room for points that fall into two boxes
    while True:
        out = torch.empty((b, cap_out, 4), dtype=torch.float32, device=device)
        counts = torch.empty((b,), dtype=torch.int32, device=device)
        if with_options:
            cabi.call("pn2_stat_rescale_opts_f64", ptr(d_raw), ptr(d_off), ptr(d_mats), ptr(d_boxes), ptr(d_boff), ptr(d_opts),
                      i32(1 if avoid_conflict else 0), ptr(rect), ptr(untouched), ptr(out), ptr(counts), ptr(box_counts),
                      ptr(box_ratio), i32(b), ctypes.c_longlong(cap), ctypes.c_longlong(cap_out), work=28.0 * float(offsets[-1]))
        else:
            cabi.call("pn2_stat_rescale_f64", ptr(d_raw), ptr(d_off), ptr(d_mats), ptr(d_boxes), ptr(d_boff), ptr(rect),
                      ptr(untouched), ptr(out), ptr(counts), ptr(box_counts), i32(b), ctypes.c_longlong(cap),
                      ctypes.c_longlong(cap_out), work=28.0 * float(offsets[-1]))
        h_counts = counts.cpu().numpy()
        if (h_counts >= 0).all():
            break
        cap_out *= 2                                             # pathological overlap of boxes: retry with more room
    h_box = box_counts.cpu().numpy()
    h_ratio = box_ratio.cpu().numpy()
    h_out = out.cpu().numpy()
    results = []
    for k in range(b):
        # norm.py:202-216: 0 for a box without points, 1 without the search, else the candidate the search stopped at
        ratios = [(RATIOS[h_ratio[j]] if avoid_conflict else 1) if h_box[j] > 0 else 0 for j in range(box_offsets[k], box_offsets[k + 1])]
        results.append((h_out[k, :int(h_counts[k])].copy(), ratios))
    return results


def convert_gpu(src, dst, spath=None, dpath=None, image_folder="image_2", calib_folder="calib", label_folder="label_2",
                use_car_sales_stats=False, avoid_conflict=False, align_front=False, rescaled_classes=("Car", "Van"),
                dataset_paths=None, batch_size=16, device=None):
    """norm.convert (norm.py:247-307) with the point rescale of `batch_size` scenes per kernel launch.  Same directory
    layout, same files (byte-identical .bin and label files)."""
    assert src in norm.datasets and dst in norm.datasets
    dataset_paths = dataset_paths or {}
    spath = spath or dataset_paths[src]
    if use_car_sales_stats:
        mapping = norm.get_scale_map(norm.car_stats_external[src], norm.car_stats_external[dst])
    else:
        mapping = norm.get_scale_map(norm.load_json(os.path.join(dataset_paths[src], "label_stats_train.json")),
                                     norm.load_json(os.path.join(dataset_paths[dst], "label_stats_train.json")))
    w, h = norm.get_image_size(spath)
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
    for start in range(0, len(names), batch_size):
        chunk = names[start:start + batch_size]
        scenes = []
        for name in chunk:
            ptc = load_velo_scan(os.path.join(spath, "training", "velodyne", "%s.bin" % name))
            calib = Calibration(os.path.join(spath, "training", calib_folder, "%s.txt" % name))
            labels = [x for x in read_label(os.path.join(spath, "training", label_folder, "%s.txt" % name))
                      if x.cls_type != "DontCare"]
            scenes.append((ptc, labels, calib))
        for name, (ptc, labels, calib), (rows, ratios) in zip(chunk, scenes,
                                                              rescale_scenes_gpu(mapping, scenes, device, rescaled_classes,
                                                                                 avoid_conflict, align_front)):
            rows.reshape(-1).tofile(os.path.join(root, "velodyne", "%s.bin" % name))
            labels = norm.scale_labels(labels, mapping, ratios, calib, w, h, align_front=align_front,
                                       rescaled_classes=rescaled_classes)
            norm.save_labels(labels, os.path.join(root, label_folder, "%s.txt" % name))
    return os.path.dirname(root)
