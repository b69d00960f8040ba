"""Mirror of evaluate/eval2.py: KITTI AP (bbox / bev / 3d / aos) with the reference's function names, argument
meaning and result layout.  The numba kernels become native code behind the C-ABI (csrc/kitti_eval.cu): the rotated
overlaps run on the GPU (rotate_iou.rotate_iou_gpu_eval = pn2_rotate_iou_eval_f32, the 3-D overlap
pn2_d3_overlap_f64), the greedy matching passes are host loops driven over whole dataset parts per call
(pn2_eval_collect_thresholds / pn2_eval_fused_statistics) instead of one numba call per image and threshold.

Pinned (tests/test_kitti_eval_*.py) against the reference module itself, imported from /root/reference in the
build container with the same IoU function substituted on both sides: result text and AP arrays equal."""
import ctypes
import io as sysio

import numpy as np

from .. import cabi
from .. import rotate_iou as _rotate_iou

rotate_iou_gpu_eval = _rotate_iou.rotate_iou_gpu_eval     # tests substitute the CPU oracle here (no GPU in CI)

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_longlong)


def _d(a):
    return a.ctypes.data_as(_f64p)


def _l(a):
    return a.ctypes.data_as(_i64p)


def _c64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(-1, cols) if cols else a


def _lib():
    lib = cabi.lib()
    if not getattr(lib, "_pn2_eval_ready", False):
        ll, dbl, i = ctypes.c_longlong, ctypes.c_double, ctypes.c_int
        lib.pn2_eval_image_box_overlap.argtypes = [_f64p, ll, _f64p, ll, i, _f64p]
        lib.pn2_eval_fused_statistics.argtypes = [_f64p, ll, ll, _f64p, _i64p, _i64p, _i64p, ll, _f64p, _f64p, _f64p, _i64p,
                                                  _i64p, i, dbl, _f64p, ll, i]
        lib.pn2_eval_collect_thresholds.argtypes = [_f64p, ll, ll, _i64p, _i64p, _i64p, ll, _f64p, _f64p, _f64p, _i64p, _i64p,
                                                    i, dbl, _f64p, _i64p]
        lib.pn2_eval_image_statistics.argtypes = [_f64p, ll, _f64p, ll, _f64p, ll, _i64p, _i64p, _f64p, ll, i, dbl, dbl, i, i,
                                                  _f64p, _f64p, _i64p]
        lib._pn2_eval_ready = True
    return lib


def get_thresholds(scores, num_gt, num_sample_pts=41):
    """eval2.py:7-25."""
    scores = np.sort(np.asarray(scores, dtype=np.float64))[::-1]
    current_recall = 0
    thresholds = []
    for i, score in enumerate(scores):
        l_recall = (i + 1) / num_gt
        if i < (len(scores) - 1):
            r_recall = (i + 2) / num_gt
        else:
            r_recall = l_recall
        if (((r_recall - current_recall) < (current_recall - l_recall)) and (i < (len(scores) - 1))):
            continue
        thresholds.append(score)
        current_recall += 1 / (num_sample_pts - 1.0)
    return thresholds


def clean_data(gt_anno, dt_anno, current_class, dataset, difficulty):
    """eval2.py:28-101: distance-band "difficulties" (0-30 / 0-70 / 0-70 / 0-30 / 30-50 / 50-70 m) with the KITTI
    occlusion / truncation limits; the height criterion is commented out in the reference and stays out."""
    CLASS_NAMES = ['car', 'pedestrian', 'cyclist']
    MAX_OCCLUSION = [0, 1, 2, 2, 2, 2]
    MAX_TRUNCATION = [0.15, 0.3, 0.5, 0.5, 0.5, 0.5]
    dc_bboxes, ignored_gt, ignored_dt = [], [], []
    current_cls_name = CLASS_NAMES[current_class].lower()
    num_gt = len(gt_anno["name"])
    num_dt = len(dt_anno["name"])
    num_valid_gt = 0
    dist_boundary = np.array([[0, 0, 0, 0, 30, 50],
                              [30, 70, 70, 30, 50, 70]])
    for i in range(num_gt):
        gt_name = gt_anno["name"][i].lower()
        if gt_name == current_cls_name:
            valid_class = 1
        elif current_cls_name == "Pedestrian".lower() and "Person_sitting".lower() == gt_name:
            valid_class = 0
        elif current_cls_name == "Car".lower() and "Van".lower() == gt_name:
            valid_class = 0
        else:
            valid_class = -1
        ignore = False
        if ((gt_anno["occluded"][i] > MAX_OCCLUSION[difficulty])
                or (gt_anno["truncated"][i] > MAX_TRUNCATION[difficulty])
                or not (dist_boundary[0, difficulty] < gt_anno["location"][i, 2] < dist_boundary[1, difficulty])):
            ignore = True
        if valid_class == 1 and not ignore:
            ignored_gt.append(0)
            num_valid_gt += 1
        elif valid_class == 0 or (ignore and (valid_class == 1)):
            ignored_gt.append(1)
        else:
            ignored_gt.append(-1)
        if gt_anno["name"][i] == "DontCare":
            dc_bboxes.append(gt_anno["bbox"][i])
    for i in range(num_dt):
        if dt_anno["name"][i].lower() == current_cls_name:
            valid_class = 1
        else:
            valid_class = -1
        if not (dist_boundary[0, difficulty] < dt_anno["location"][i, 2] < dist_boundary[1, difficulty]):
            ignored_dt.append(1)
        elif valid_class == 1:
            ignored_dt.append(0)
        else:
            ignored_dt.append(-1)
    return num_valid_gt, ignored_gt, ignored_dt, dc_bboxes


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """eval2.py:104-127 (float64)."""
    boxes, query_boxes = _c64(boxes, 4), _c64(query_boxes, 4)
    out = np.zeros((boxes.shape[0], query_boxes.shape[0]), dtype=np.float64)
    cabi.check(_lib().pn2_eval_image_box_overlap(_d(boxes), boxes.shape[0], _d(query_boxes), query_boxes.shape[0],
                                                 int(criterion), _d(out)), "pn2_eval_image_box_overlap")
    return out


def bev_box_overlap(boxes, qboxes, criterion=-1):
    """eval2.py:130-132."""
    return rotate_iou_gpu_eval(boxes, qboxes, criterion)


def d3_box_overlap_kernel(boxes, qboxes, rinc, criterion=-1):
    """eval2.py:136-162, in place on rinc like the numba loop; on the device when one is present (pn2_d3_overlap_f64),
    else the same float64 expressions in numpy (CPU-only unit tests)."""
    import torch
    boxes, qboxes = _c64(boxes, 7), _c64(qboxes, 7)
    n, k = boxes.shape[0], qboxes.shape[0]
    if n == 0 or k == 0:
        return
    if torch.cuda.is_available():
        from ..cabi import i32, ptr
        db, dq = torch.from_numpy(boxes).cuda(), torch.from_numpy(qboxes).cuda()
        dr = torch.from_numpy(np.ascontiguousarray(rinc, dtype=np.float32)).cuda()
        out = torch.empty((n, k), dtype=torch.float64, device=db.device)
        cabi.call("pn2_d3_overlap_f64", ptr(db), ctypes.c_longlong(n), ptr(dq), ctypes.c_longlong(k), ptr(dr),
                  i32(criterion), ptr(out), work=float(n) * k)
        rinc[...] = out.cpu().numpy()
        return
    r = np.asarray(rinc, dtype=np.float64)
    iw = np.minimum(boxes[:, None, 1], qboxes[None, :, 1]) - np.maximum(boxes[:, None, 1] - boxes[:, None, 4],
                                                                         qboxes[None, :, 1] - qboxes[None, :, 4])
    area1 = (boxes[:, 3] * boxes[:, 4] * boxes[:, 5])[:, None]
    area2 = (qboxes[:, 3] * qboxes[:, 4] * qboxes[:, 5])[None, :]
    inc = iw * r
    if criterion == -1:
        ua = area1 + area2 - inc
    elif criterion == 0:
        ua = area1 + 0 * inc
    elif criterion == 1:
        ua = area2 + 0 * inc
    else:
        ua = inc
    with np.errstate(divide='ignore', invalid='ignore'):
        val = np.where(iw > 0, inc / ua, 0.0)
    rinc[...] = np.where(r > 0, val, r)


def d3_box_overlap(boxes, qboxes, criterion=-1):
    """eval2.py:165-169."""
    rinc = rotate_iou_gpu_eval(boxes[:, [0, 2, 3, 5, 6]], qboxes[:, [0, 2, 3, 5, 6]], 2)
    # float32 like the reference's: its in-place numba loop rounds every inc / ua to rinc's dtype before the match
    # thresholds see it
    rinc = np.ascontiguousarray(rinc, dtype=np.float32)
    d3_box_overlap_kernel(boxes, qboxes, rinc, criterion)
    return rinc


def get_split_parts(num, num_part):
    """eval2.py:301-308."""
    same_part = num // num_part
    remain_num = num % num_part
    if remain_num == 0:
        return [same_part] * num_part
    return [same_part] * num_part + [remain_num]


def compute_statistics_jit(overlaps, gt_datas, dt_datas, ignored_gt, ignored_det, dc_bboxes, metric, min_overlap,
                           thresh=0, compute_fp=False, compute_aos=False):
    """eval2.py:172-298 for ONE image -> (tp, fp, fn, similarity, thresholds), the reference's values in every mode
    (similarity -1 where it is undefined).  eval_class drives whole parts per native call instead."""
    ov = _c64(overlaps)
    gt, dt, dc = _c64(gt_datas, 5), _c64(dt_datas, 6), _c64(dc_bboxes, 4)
    ig, idt = np.ascontiguousarray(ignored_gt, np.int64), np.ascontiguousarray(ignored_det, np.int64)
    out4 = np.zeros((4,), np.float64)
    th = np.zeros((max(gt.shape[0], 1),), np.float64)
    n = np.zeros((1,), np.int64)
    cabi.check(_lib().pn2_eval_image_statistics(_d(ov), gt.shape[0], _d(gt), gt.shape[0], _d(dt), dt.shape[0], _l(ig),
                                                _l(idt), _d(dc), dc.shape[0], int(metric), float(min_overlap),
                                                float(thresh), 1 if compute_fp else 0, 1 if compute_aos else 0, _d(out4),
                                                _d(th), _l(n)), "pn2_eval_image_statistics")
    sim = out4[3] if (compute_fp and compute_aos) else 0
    return int(out4[0]), int(out4[1]), int(out4[2]), (float(sim) if sim != -1 else -1), th[:int(n[0])]


def fused_compute_statistics(overlaps, pr, gt_nums, dt_nums, dc_nums, gt_datas, dt_datas, dontcares, ignored_gts,
                             ignored_dets, metric, min_overlap, thresholds, compute_aos=False):
    """eval2.py:311-358: pr (n_thresholds, 4) += [tp, fp, fn, similarity] over the images of one part."""
    ov = _c64(overlaps)
    th = _c64(thresholds)
    assert pr.dtype == np.float64 and pr.flags.c_contiguous
    cabi.check(_lib().pn2_eval_fused_statistics(
        _d(ov), ov.shape[0], ov.shape[1] if ov.ndim == 2 else 0, _d(pr), _l(np.ascontiguousarray(gt_nums, np.int64)),
        _l(np.ascontiguousarray(dt_nums, np.int64)), _l(np.ascontiguousarray(dc_nums, np.int64)), len(gt_nums),
        _d(_c64(gt_datas, 5)), _d(_c64(dt_datas, 6)), _d(_c64(dontcares, 4)), _l(np.ascontiguousarray(ignored_gts, np.int64)),
        _l(np.ascontiguousarray(ignored_dets, np.int64)), int(metric), float(min_overlap), _d(th), th.shape[0],
        1 if compute_aos else 0), "pn2_eval_fused_statistics")


def calculate_iou_partly(gt_annos, dt_annos, metric, num_parts=50):
    """eval2.py:361-432 (camera coordinates; metric 0: bbox, 1: bev, 2: 3d)."""
    assert len(gt_annos) == len(dt_annos)
    total_dt_num = np.stack([len(a["name"]) for a in dt_annos], 0)
    total_gt_num = np.stack([len(a["name"]) for a in gt_annos], 0)
    num_examples = len(gt_annos)
    split_parts = get_split_parts(num_examples, num_parts)
    parted_overlaps = []
    example_idx = 0

    def boxes_of(annos, cols):
        loc = np.concatenate([a["location"][:, cols] if cols else a["location"] for a in annos], 0)
        dims = np.concatenate([a["dimensions"][:, cols] if cols else a["dimensions"] for a in annos], 0)
        rots = np.concatenate([a["rotation_y"] for a in annos], 0)
        return np.concatenate([loc, dims, rots[..., np.newaxis]], axis=1)

    for num_part in split_parts:
        gt_annos_part = gt_annos[example_idx:example_idx + num_part]
        dt_annos_part = dt_annos[example_idx:example_idx + num_part]
        if metric == 0:
            gt_boxes = np.concatenate([a["bbox"] for a in gt_annos_part], 0)
            dt_boxes = np.concatenate([a["bbox"] for a in dt_annos_part], 0)
            overlap_part = image_box_overlap(gt_boxes, dt_boxes)
        elif metric == 1:
            overlap_part = bev_box_overlap(boxes_of(gt_annos_part, [0, 2]), boxes_of(dt_annos_part, [0, 2])).astype(np.float64)
        elif metric == 2:
            overlap_part = d3_box_overlap(boxes_of(gt_annos_part, None), boxes_of(dt_annos_part, None)).astype(np.float64)
        else:
            raise ValueError("unknown metric")
        parted_overlaps.append(overlap_part)
        example_idx += num_part
    overlaps = []
    example_idx = 0
    for j, num_part in enumerate(split_parts):
        gt_num_idx, dt_num_idx = 0, 0
        for i in range(num_part):
            gt_box_num = total_gt_num[example_idx + i]
            dt_box_num = total_dt_num[example_idx + i]
            overlaps.append(parted_overlaps[j][gt_num_idx:gt_num_idx + gt_box_num, dt_num_idx:dt_num_idx + dt_box_num])
            gt_num_idx += gt_box_num
            dt_num_idx += dt_box_num
        example_idx += num_part
    return overlaps, parted_overlaps, total_gt_num, total_dt_num


def _prepare_data(gt_annos, dt_annos, current_class, dataset, difficulty):
    """eval2.py:435-464."""
    gt_datas_list, dt_datas_list, total_dc_num = [], [], []
    ignored_gts, ignored_dets, dontcares = [], [], []
    total_num_valid_gt = 0
    for i in range(len(gt_annos)):
        num_valid_gt, ignored_gt, ignored_det, dc_bboxes = clean_data(gt_annos[i], dt_annos[i], current_class, dataset, difficulty)
        ignored_gts.append(np.array(ignored_gt, dtype=np.int64))
        ignored_dets.append(np.array(ignored_det, dtype=np.int64))
        if len(dc_bboxes) == 0:
            dc_bboxes = np.zeros((0, 4)).astype(np.float64)
        else:
            dc_bboxes = np.stack(dc_bboxes, 0).astype(np.float64)
        total_dc_num.append(dc_bboxes.shape[0])
        dontcares.append(dc_bboxes)
        total_num_valid_gt += num_valid_gt
        gt_datas = np.concatenate([gt_annos[i]["bbox"], gt_annos[i]["alpha"][..., np.newaxis]], 1)
        dt_datas = np.concatenate([dt_annos[i]["bbox"], dt_annos[i]["alpha"][..., np.newaxis],
                                   dt_annos[i]["score"][..., np.newaxis]], 1)
        gt_datas_list.append(gt_datas)
        dt_datas_list.append(dt_datas)
    total_dc_num = np.stack(total_dc_num, axis=0)
    print(f"difficulty: {difficulty}, total_num_valid_gt: {total_num_valid_gt}")
    return (gt_datas_list, dt_datas_list, ignored_gts, ignored_dets, dontcares, total_dc_num, total_num_valid_gt)


def eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, metric, min_overlaps, compute_aos=False,
               num_parts=50):
    """eval2.py:467-563 -> {"recall", "precision", "orientation"} of shape [class, difficulty, min_overlap, 41].
    The two statistics passes run part-wise in native code (one call per part instead of one numba call per image
    and per threshold); everything else is the reference's flow."""
    assert len(gt_annos) == len(dt_annos)
    num_examples = len(gt_annos)
    split_parts = get_split_parts(num_examples, num_parts)
    rets = calculate_iou_partly(dt_annos, gt_annos, metric, num_parts)
    overlaps, parted_overlaps, total_dt_num, total_gt_num = rets
    N_SAMPLE_PTS = 41
    num_minoverlap = len(min_overlaps)
    num_class = len(current_classes)
    num_difficulty = len(difficultys)
    precision = np.zeros([num_class, num_difficulty, num_minoverlap, N_SAMPLE_PTS])
    recall = np.zeros([num_class, num_difficulty, num_minoverlap, N_SAMPLE_PTS])
    aos = np.zeros([num_class, num_difficulty, num_minoverlap, N_SAMPLE_PTS])
    lib = _lib()
    parted = [_c64(p) for p in parted_overlaps]
    for m, current_class in enumerate(current_classes):
        for l, difficulty in enumerate(difficultys):
            rets = _prepare_data(gt_annos, dt_annos, current_class, dataset, difficulty)
            (gt_datas_list, dt_datas_list, ignored_gts, ignored_dets, dontcares, total_dc_num, total_num_valid_gt) = rets
            # the per-part concatenations of eval2.py:524-533, built once per (class, difficulty)
            parts = []
            idx = 0
            for j, num_part in enumerate(split_parts):
                sl = slice(idx, idx + num_part)
                cat = lambda xs, cols, dt: (np.ascontiguousarray(np.concatenate(xs, 0), dtype=dt).reshape(-1, cols) if cols
                                            else np.ascontiguousarray(np.concatenate(xs, 0), dtype=dt)) if len(xs) else \
                    np.zeros((0, cols) if cols else (0,), dt)
                parts.append(dict(
                    ov=parted[j], gt=cat(gt_datas_list[sl], 5, np.float64), dt=cat(dt_datas_list[sl], 6, np.float64),
                    dc=cat(dontcares[sl], 4, np.float64), ig=cat(ignored_gts[sl], 0, np.int64),
                    idt=cat(ignored_dets[sl], 0, np.int64), gn=np.ascontiguousarray(total_gt_num[sl], np.int64),
                    dn=np.ascontiguousarray(total_dt_num[sl], np.int64), dcn=np.ascontiguousarray(total_dc_num[sl], np.int64)))
                idx += num_part
            for k, min_overlap in enumerate(min_overlaps[:, metric, m]):
                thresholdss = []
                for p in parts:                                             # eval2.py:506-520, part-wise
                    if len(p["gn"]) == 0:
                        continue
                    th = np.zeros((max(int(p["gn"].sum()), 1),), np.float64)
                    n = ctypes.c_longlong(0)
                    cabi.check(lib.pn2_eval_collect_thresholds(
                        _d(p["ov"]), p["ov"].shape[0], p["ov"].shape[1], _l(p["gn"]), _l(p["dn"]), _l(p["dcn"]), len(p["gn"]),
                        _d(p["gt"]), _d(p["dt"]), _d(p["dc"]), _l(p["ig"]), _l(p["idt"]), int(metric), float(min_overlap),
                        _d(th), ctypes.byref(n)), "pn2_eval_collect_thresholds")
                    thresholdss += th[:n.value].tolist()
                thresholdss = np.array(thresholdss)
                thresholds = get_thresholds(thresholdss, total_num_valid_gt)
                thresholds = np.array(thresholds, dtype=np.float64)
                pr = np.zeros([len(thresholds), 4])
                for p in parts:                                             # eval2.py:522-550
                    if len(p["gn"]) == 0 or len(thresholds) == 0:
                        continue
                    cabi.check(lib.pn2_eval_fused_statistics(
                        _d(p["ov"]), p["ov"].shape[0], p["ov"].shape[1], _d(pr), _l(p["gn"]), _l(p["dn"]), _l(p["dcn"]),
                        len(p["gn"]), _d(p["gt"]), _d(p["dt"]), _d(p["dc"]), _l(p["ig"]), _l(p["idt"]), int(metric),
                        float(min_overlap), _d(thresholds), len(thresholds), 1 if compute_aos else 0),
                        "pn2_eval_fused_statistics")
                with np.errstate(divide='ignore', invalid='ignore'):
                    for i in range(len(thresholds)):
                        recall[m, l, k, i] = pr[i, 0] / (pr[i, 0] + pr[i, 2])
                        precision[m, l, k, i] = pr[i, 0] / (pr[i, 0] + pr[i, 1])
                        if compute_aos:
                            aos[m, l, k, i] = pr[i, 3] / (pr[i, 0] + pr[i, 1])
                for i in range(len(thresholds)):
                    precision[m, l, k, i] = np.max(precision[m, l, k, i:], axis=-1)
                    recall[m, l, k, i] = np.max(recall[m, l, k, i:], axis=-1)
                    if compute_aos:
                        aos[m, l, k, i] = np.max(aos[m, l, k, i:], axis=-1)
    return {"recall": recall, "precision": precision, "orientation": aos}


def get_mAP(prec):
    """eval2.py:566-570: 11-point interpolation over the 41 recall samples."""
    sums = 0
    for i in range(0, prec.shape[-1], 4):
        sums = sums + prec[..., i]
    return sums / 11 * 100


def print_str(value, *arg, sstream=None):
    if sstream is None:
        sstream = sysio.StringIO()
    sstream.truncate(0)
    sstream.seek(0)
    print(value, *arg, file=sstream)
    return sstream.getvalue()


def do_eval(gt_annos, dt_annos, current_classes, dataset, min_overlaps, compute_aos=False):
    """eval2.py:583-603.  min_overlaps: [num_minoverlap, metric, num_class]."""
    difficultys = [0, 1, 2, 3, 4, 5]
    ret = eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, 0, min_overlaps, compute_aos)
    mAP_bbox = get_mAP(ret["precision"])
    mAP_aos = None
    if compute_aos:
        mAP_aos = get_mAP(ret["orientation"])
    ret = eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, 1, min_overlaps)
    mAP_bev = get_mAP(ret["precision"])
    ret = eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, 2, min_overlaps)
    mAP_3d = get_mAP(ret["precision"])
    return mAP_bbox, mAP_bev, mAP_3d, mAP_aos


def get_official_eval_result(gt_annos, dt_annos, current_classes, dataset, dense_sample=False):
    """eval2.py:624-710 -> (result text, dict)."""
    overlap_0_7 = np.array([[0.7, 0.5, 0.5, 0.7, 0.5], [0.7, 0.5, 0.5, 0.7, 0.5], [0.7, 0.5, 0.5, 0.7, 0.5]])
    overlap_0_5 = np.array([[0.7, 0.5, 0.5, 0.7, 0.5], [0.5, 0.25, 0.25, 0.5, 0.25], [0.5, 0.25, 0.25, 0.5, 0.25]])
    overlaps = []
    if dense_sample:
        for i in range(101):
            tmp = np.zeros((3, 5))
            tmp[:, 0] = i / 100.0
            overlaps.append(tmp)
    min_overlaps = np.stack([overlap_0_7, overlap_0_5] + overlaps, axis=0)
    class_to_name = {0: 'Car', 1: 'Pedestrian', 2: 'Cyclist', 3: 'Van', 4: 'Person_sitting'}
    name_to_class = {v: n for n, v in class_to_name.items()}
    if not isinstance(current_classes, (list, tuple)):
        current_classes = [current_classes]
    current_classes = [name_to_class[c] if isinstance(c, str) else c for c in current_classes]
    min_overlaps = min_overlaps[:, :, current_classes]
    result = ''
    compute_aos = False
    for anno in dt_annos:
        if anno['alpha'].shape[0] != 0:
            if anno['alpha'][0] != -10:
                compute_aos = True
            break
    mAPbbox, mAPbev, mAP3d, mAPaos = do_eval(gt_annos, dt_annos, current_classes, dataset, min_overlaps, compute_aos)
    ret_dict = {}
    res = dict()
    for j, curcls in enumerate(current_classes):
        res[curcls] = dict()
        for i in range(min_overlaps.shape[0]):
            key = f"{class_to_name[curcls]} " + "AP@{:.2f}, {:.2f}, {:.2f}".format(*min_overlaps[i, :, j])
            res[curcls][key] = dict()
            res[curcls][key]["mAPbbox"] = mAPbbox[j, :, i]
            res[curcls][key]["mAPbev"] = mAPbev[j, :, i]
            res[curcls][key]["mAP3d"] = mAP3d[j, :, i]
            result += print_str((f"{class_to_name[curcls]} " + "AP@{:.2f}, {:.2f}, {:.2f}:".format(*min_overlaps[i, :, j])))
            for tag, arr in (("bbox", mAPbbox), ("bev ", mAPbev), ("3d  ", mAP3d)):
                result += print_str(f"{tag} AP:" + "".join(f"{arr[j, d, i]:.4f}, " for d in range(6)))
            if compute_aos:
                result += print_str("aos  AP:" + ", ".join(f"{mAPaos[j, d, i]:.2f}" for d in range(6)))
    ret_dict['Car_3d_easy'] = mAP3d[0, 0, 0]
    ret_dict['Car_3d_moderate'] = mAP3d[0, 1, 0]
    ret_dict['Car_3d_hard'] = mAP3d[0, 2, 0]
    ret_dict['Car_bev_easy'] = mAPbev[0, 0, 0]
    ret_dict['Car_bev_moderate'] = mAPbev[0, 1, 0]
    ret_dict['Car_bev_hard'] = mAPbev[0, 2, 0]
    ret_dict['Car_image_easy'] = mAPbbox[0, 0, 0]
    ret_dict['Car_image_moderate'] = mAPbbox[0, 1, 0]
    ret_dict['Car_image_hard'] = mAPbbox[0, 2, 0]
    ret_dict["result"] = res
    return result, ret_dict
