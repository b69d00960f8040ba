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


N_RECALL_SAMPLES = 41
# one row per "difficulty" of this fork: distance bands in metres (camera z) with KITTI's occlusion / truncation limits
# (eval2.py:31-41: 0-30 easy, 0-70 moderate, 0-70 hard, then the 0-30 / 30-50 / 50-70 m range splits)
DIFFICULTY_TABLE = (
    # (z_low, z_high, max_occlusion, max_truncation)
    (0, 30, 0, 0.15),
    (0, 70, 1, 0.3),
    (0, 70, 2, 0.5),
    (0, 30, 2, 0.5),
    (30, 50, 2, 0.5),
    (50, 70, 2, 0.5),
)
EVAL_CLASS_NAMES = ('car', 'pedestrian', 'cyclist')
# ground-truth classes that are neither a hit nor a miss for a detector of the key class (eval2.py:48-53)
NEIGHBOUR_CLASS = {'pedestrian': 'person_sitting', 'car': 'van'}


def get_thresholds(scores, num_gt, num_sample_pts=N_RECALL_SAMPLES):
    """eval2.py:7-25: the score thresholds at which recall crosses the num_sample_pts equally spaced sample positions.
    Same arithmetic (rank / num_gt recalls, a running target advanced by 1 / (num_sample_pts - 1.0))."""
    ranked = np.sort(np.asarray(scores, dtype=np.float64))[::-1]
    step = 1 / (num_sample_pts - 1.0)
    target = 0
    picked = []
    final = len(ranked) - 1
    for pos, score in enumerate(ranked):
        here = (pos + 1) / num_gt
        beyond = (pos + 2) / num_gt if pos < final else here
        if pos < final and (beyond - target) < (target - here):
            continue                                   # the next detection gets closer to the target recall
        picked.append(score)
        target += step
    return picked


def _outside_band(z, difficulty):
    low, high = DIFFICULTY_TABLE[difficulty][0], DIFFICULTY_TABLE[difficulty][1]
    return ~((low < z) & (z < high))                   # NaN depths are outside, like the reference's chained compare


def clean_data(gt_anno, dt_anno, current_class, dataset, difficulty):
    """eval2.py:28-101 -> (number of valid ground truths, ignored_gt, ignored_dt, DontCare boxes).
    ignored_*: 0 = counts, 1 = neither hit nor miss (neighbour class, too occluded / truncated, outside the distance
    band), -1 = another class.  The height criterion is commented out in the reference and stays out.  Vectorised
    over the boxes of the image; `dataset` is accepted for signature compatibility and unused, as upstream."""
    key = EVAL_CLASS_NAMES[current_class]               # IndexError beyond the three evaluated classes, as upstream
    neighbour = NEIGHBOUR_CLASS.get(key)
    _, _, max_occlusion, max_truncation = DIFFICULTY_TABLE[difficulty]

    gt_names = [str(n).lower() for n in gt_anno["name"]]
    same = np.array([n == key for n in gt_names], dtype=bool)
    near_class = np.array([n == neighbour for n in gt_names], dtype=bool)
    if len(gt_names):
        too_hard = ((np.asarray(gt_anno["occluded"]) > max_occlusion) | (np.asarray(gt_anno["truncated"]) > max_truncation)
                    | _outside_band(np.asarray(gt_anno["location"])[:, 2], difficulty))
    else:
        too_hard = np.zeros((0,), dtype=bool)
    counted = same & ~too_hard
    ignored_gt = np.where(counted, 0, np.where(near_class | same, 1, -1))
    dc_bboxes = [gt_anno["bbox"][i] for i, n in enumerate(gt_anno["name"]) if n == "DontCare"]

    dt_names = [str(n).lower() for n in dt_anno["name"]]
    if len(dt_names):
        dt_same = np.array([n == key for n in dt_names], dtype=bool)
        ignored_dt = np.where(_outside_band(np.asarray(dt_anno["location"])[:, 2], difficulty), 1, np.where(dt_same, 0, -1))
    else:
        ignored_dt = np.zeros((0,), dtype=np.int64)
    return int(counted.sum()), ignored_gt.astype(np.int64).tolist(), ignored_dt.astype(np.int64).tolist(), dc_bboxes


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
    """eval2.py:301-308: num_part equal chunks of num // num_part images, plus one chunk with the remainder."""
    chunk, rest = divmod(num, num_part)
    return [chunk] * num_part + ([rest] if rest else [])


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


def _part_bounds(num_images, num_parts):
    """[(first image, one past the last image)] of every dataset part (get_split_parts as index ranges)"""
    edges = np.concatenate(([0], np.cumsum(get_split_parts(num_images, num_parts)))).astype(np.int64)
    return list(zip(edges[:-1].tolist(), edges[1:].tolist()))


def _camera_boxes(annos, cols):
    """[location | dimensions | rotation_y] rows of all boxes of `annos`; cols picks the (x, z) / (l, w) pair for BEV"""
    pick = (lambda a: a[:, cols]) if cols is not None else (lambda a: a)
    return np.concatenate([np.concatenate([pick(a["location"]) for a in annos], 0),
                           np.concatenate([pick(a["dimensions"]) for a in annos], 0),
                           np.concatenate([a["rotation_y"] for a in annos], 0)[..., np.newaxis]], axis=1)


def _part_overlaps(first_annos, second_annos, metric):
    """all-pairs overlap matrix (boxes of first_annos x boxes of second_annos) of one dataset part"""
    if metric == 0:
        return image_box_overlap(np.concatenate([a["bbox"] for a in first_annos], 0),
                                 np.concatenate([a["bbox"] for a in second_annos], 0))
    if metric == 1:
        return bev_box_overlap(_camera_boxes(first_annos, [0, 2]), _camera_boxes(second_annos, [0, 2])).astype(np.float64)
    if metric == 2:
        return d3_box_overlap(_camera_boxes(first_annos, None), _camera_boxes(second_annos, None)).astype(np.float64)
    raise ValueError("unknown metric")


def calculate_iou_partly(gt_annos, dt_annos, metric, num_parts=50):
    """eval2.py:361-432 (camera coordinates; metric 0: bbox, 1: bev, 2: 3d) -> (per-image overlap blocks, per-part
    matrices, boxes per image of the first / of the second argument).  One overlap call per part; the per-image blocks
    are views on the diagonal of the part matrix."""
    assert len(gt_annos) == len(dt_annos)
    first_counts = np.stack([len(a["name"]) for a in gt_annos], 0)
    second_counts = np.stack([len(a["name"]) for a in dt_annos], 0)
    per_image, per_part = [], []
    for lo, hi in _part_bounds(len(gt_annos), num_parts):
        matrix = _part_overlaps(gt_annos[lo:hi], dt_annos[lo:hi], metric)
        per_part.append(matrix)
        row = col = 0
        for image in range(lo, hi):
            rows, cols = first_counts[image], second_counts[image]
            per_image.append(matrix[row:row + rows, col:col + cols])
            row, col = row + rows, col + cols
    return per_image, per_part, first_counts, second_counts


def _prepare_data(gt_annos, dt_annos, current_class, dataset, difficulty):
    """eval2.py:435-464: per image [bbox | alpha] of the ground truths, [bbox | alpha | score] of the detections, the
    ignore codes of clean_data, the DontCare boxes and their number; the count of valid ground truths of the data set."""
    gt_rows, dt_rows, gt_codes, dt_codes, dc_boxes = [], [], [], [], []
    valid_total = 0
    for gt, dt in zip(gt_annos, dt_annos):
        n_valid, code_gt, code_dt, dc = clean_data(gt, dt, current_class, dataset, difficulty)
        valid_total += n_valid
        gt_codes.append(np.array(code_gt, dtype=np.int64))
        dt_codes.append(np.array(code_dt, dtype=np.int64))
        dc_boxes.append(np.stack(dc, 0).astype(np.float64) if len(dc) else np.zeros((0, 4), dtype=np.float64))
        gt_rows.append(np.concatenate([gt["bbox"], gt["alpha"][..., np.newaxis]], 1))
        dt_rows.append(np.concatenate([dt["bbox"], dt["alpha"][..., np.newaxis], dt["score"][..., np.newaxis]], 1))
    dc_counts = np.stack([b.shape[0] for b in dc_boxes], axis=0)
    print(f"difficulty: {difficulty}, total_num_valid_gt: {valid_total}")
    return gt_rows, dt_rows, gt_codes, dt_codes, dc_boxes, dc_counts, valid_total


class _Part:
    """One dataset part packed for the native matching passes (csrc/kitti_eval.cu): the overlap matrix of the part and the
    concatenated per-image arrays of eval2.py:524-533, built once per (class, difficulty) and reused by every threshold."""

    @staticmethod
    def _cat(arrays, cols, dtype):
        if not len(arrays):
            return np.zeros((0, cols) if cols else (0,), dtype)
        joined = np.ascontiguousarray(np.concatenate(arrays, 0), dtype=dtype)
        return joined.reshape(-1, cols) if cols else joined

    def __init__(self, matrix, images, prepared, first_counts, second_counts):
        gt_rows, dt_rows, gt_codes, dt_codes, dc_boxes, dc_counts, _ = prepared
        self.overlaps = matrix
        self.gt = self._cat(gt_rows[images], 5, np.float64)
        self.dt = self._cat(dt_rows[images], 6, np.float64)
        self.dc = self._cat(dc_boxes[images], 4, np.float64)
        self.gt_codes = self._cat(gt_codes[images], 0, np.int64)
        self.dt_codes = self._cat(dt_codes[images], 0, np.int64)
        # eval_class computes overlaps(dt, gt): the matrix rows are detections, its columns ground truths
        self.n_gt = np.ascontiguousarray(second_counts[images], np.int64)
        self.n_dt = np.ascontiguousarray(first_counts[images], np.int64)
        self.n_dc = np.ascontiguousarray(dc_counts[images], np.int64)
        self.images = len(self.n_gt)

    def _common(self):
        return (_d(self.overlaps), self.overlaps.shape[0], self.overlaps.shape[1])

    def true_positive_scores(self, lib, metric, min_overlap):
        """compute_statistics_jit(thresh = 0, compute_fp = False) over the part's images (eval2.py:506-520)"""
        scores = np.zeros((max(int(self.n_gt.sum()), 1),), np.float64)
        count = ctypes.c_longlong(0)
        cabi.check(lib.pn2_eval_collect_thresholds(
            *self._common(), _l(self.n_gt), _l(self.n_dt), _l(self.n_dc), self.images, _d(self.gt), _d(self.dt), _d(self.dc),
            _l(self.gt_codes), _l(self.dt_codes), int(metric), float(min_overlap), _d(scores), ctypes.byref(count)),
            "pn2_eval_collect_thresholds")
        return scores[:count.value]

    def accumulate(self, lib, pr, metric, min_overlap, thresholds, compute_aos):
        """fused_compute_statistics (eval2.py:522-550): pr (n_thresholds, 4) += [tp, fp, fn, similarity]"""
        cabi.check(lib.pn2_eval_fused_statistics(
            *self._common(), _d(pr), _l(self.n_gt), _l(self.n_dt), _l(self.n_dc), self.images, _d(self.gt), _d(self.dt),
            _d(self.dc), _l(self.gt_codes), _l(self.dt_codes), int(metric), float(min_overlap), _d(thresholds),
            len(thresholds), 1 if compute_aos else 0), "pn2_eval_fused_statistics")


def _running_max_from_the_right(curve, n):
    """curve[i] = max(curve[i:]) for i < n, in place (eval2.py:557-562; curve has 41 entries, n of them are filled)"""
    for i in range(n):
        curve[i] = np.max(curve[i:], axis=-1)


def eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, metric, min_overlaps, compute_aos=False,
               num_parts=50):
    """eval2.py:467-563 -> {"recall", "precision", "orientation"} of shape [class, difficulty, min_overlap, 41].
    The two statistics passes run part-wise in native code (one call per part instead of one numba call per image and
    per threshold)."""
    assert len(gt_annos) == len(dt_annos)
    bounds = _part_bounds(len(gt_annos), num_parts)
    # argument order as upstream (eval2.py:492): detections first
    _, matrices, first_counts, second_counts = calculate_iou_partly(dt_annos, gt_annos, metric, num_parts)
    matrices = [_c64(m) for m in matrices]
    shape = [len(current_classes), len(difficultys), len(min_overlaps), N_RECALL_SAMPLES]
    precision, recall, aos = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    lib = _lib()
    for m, current_class in enumerate(current_classes):
        for l, difficulty in enumerate(difficultys):
            prepared = _prepare_data(gt_annos, dt_annos, current_class, dataset, difficulty)
            valid_total = prepared[-1]
            parts = [_Part(matrices[j], slice(lo, hi), prepared, first_counts, second_counts)
                     for j, (lo, hi) in enumerate(bounds)]
            parts = [part for part in parts if part.images]
            for k, min_overlap in enumerate(min_overlaps[:, metric, m]):
                scores = [part.true_positive_scores(lib, metric, min_overlap) for part in parts]
                scores = np.concatenate(scores) if scores else np.zeros((0,))
                thresholds = np.array(get_thresholds(scores, valid_total), dtype=np.float64)
                n = len(thresholds)
                pr = np.zeros([n, 4])
                if n:
                    for part in parts:
                        part.accumulate(lib, pr, metric, min_overlap, thresholds, compute_aos)
                tp, fp, fn, similarity = pr[:, 0], pr[:, 1], pr[:, 2], pr[:, 3]
                with np.errstate(divide='ignore', invalid='ignore'):
                    recall[m, l, k, :n] = tp / (tp + fn)
                    precision[m, l, k, :n] = tp / (tp + fp)
                    if compute_aos:
                        aos[m, l, k, :n] = similarity / (tp + fp)
                _running_max_from_the_right(precision[m, l, k], n)
                _running_max_from_the_right(recall[m, l, k], n)
                if compute_aos:
                    _running_max_from_the_right(aos[m, l, k], n)
    return {"recall": recall, "precision": precision, "orientation": aos}


def get_mAP(prec):
    """eval2.py:566-570: 11-point interpolation, every fourth of the 41 recall samples (summed in index order)."""
    total = 0
    for sample in range(0, prec.shape[-1], 4):
        total = total + prec[..., sample]
    return total / 11 * 100


def print_str(value, *arg, sstream=None):
    """eval2.py:573-579: what print() would write, as a string"""
    stream = sysio.StringIO() if sstream is None else sstream
    stream.truncate(0)
    stream.seek(0)
    print(value, *arg, file=stream)
    return stream.getvalue()


def do_eval(gt_annos, dt_annos, current_classes, dataset, min_overlaps, compute_aos=False):
    """eval2.py:583-603 -> (mAP bbox, bev, 3d, aos | None), each [class, difficulty, min_overlap].
    min_overlaps: [num_minoverlap, metric, num_class]."""
    difficultys = list(range(len(DIFFICULTY_TABLE)))
    curves = {}
    for metric, tag in enumerate(("bbox", "bev", "3d")):
        curves[tag] = eval_class(gt_annos, dt_annos, current_classes, dataset, difficultys, metric, min_overlaps,
                                 compute_aos and metric == 0)
    aos = get_mAP(curves["bbox"]["orientation"]) if compute_aos else None
    return get_mAP(curves["bbox"]["precision"]), get_mAP(curves["bev"]["precision"]), get_mAP(curves["3d"]["precision"]), aos


CLASS_TO_NAME = {0: 'Car', 1: 'Pedestrian', 2: 'Cyclist', 3: 'Van', 4: 'Person_sitting'}
# [metric bbox / bev / 3d][class]: KITTI's official overlaps, and the relaxed set (eval2.py:625-630)
OVERLAP_STRICT = ((0.7, 0.5, 0.5, 0.7, 0.5),) * 3
OVERLAP_RELAXED = ((0.7, 0.5, 0.5, 0.7, 0.5), (0.5, 0.25, 0.25, 0.5, 0.25), (0.5, 0.25, 0.25, 0.5, 0.25))


def _detections_carry_alpha(dt_annos):
    """orientation similarity is evaluated when the first non-empty detection set has a real alpha (eval2.py:658-664)"""
    for anno in dt_annos:
        if anno['alpha'].shape[0] != 0:
            return bool(anno['alpha'][0] != -10)
    return False


def get_official_eval_result(gt_annos, dt_annos, current_classes, dataset, dense_sample=False):
    """eval2.py:624-710 -> (result text, dict).  Text layout: per class and overlap set a header line, then one line per
    metric with the six distance-band APs (%.4f, trailing ", "), then the AOS line (%.2f) when alpha is present."""
    sets = [np.array(OVERLAP_STRICT), np.array(OVERLAP_RELAXED)]
    if dense_sample:
        for percent in range(101):
            dense = np.zeros((3, 5))
            dense[:, 0] = percent / 100.0
            sets.append(dense)
    name_to_class = {name: index for index, name in CLASS_TO_NAME.items()}
    if not isinstance(current_classes, (list, tuple)):
        current_classes = [current_classes]
    current_classes = [name_to_class[c] if isinstance(c, str) else c for c in current_classes]
    min_overlaps = np.stack(sets, axis=0)[:, :, current_classes]                 # [overlap set, metric, class]
    compute_aos = _detections_carry_alpha(dt_annos)
    ap_bbox, ap_bev, ap_3d, ap_aos = do_eval(gt_annos, dt_annos, current_classes, dataset, min_overlaps, compute_aos)

    lines = []
    per_class = {}
    for j, cls in enumerate(current_classes):
        per_class[cls] = {}
        for i in range(min_overlaps.shape[0]):
            title = f"{CLASS_TO_NAME[cls]} " + "AP@{:.2f}, {:.2f}, {:.2f}".format(*min_overlaps[i, :, j])
            per_class[cls][title] = {"mAPbbox": ap_bbox[j, :, i], "mAPbev": ap_bev[j, :, i], "mAP3d": ap_3d[j, :, i]}
            lines.append(title + ":")
            for label, table in (("bbox AP:", ap_bbox), ("bev  AP:", ap_bev), ("3d   AP:", ap_3d)):
                lines.append(label + "".join(f"{value:.4f}, " for value in table[j, :6, i]))
            if compute_aos:
                lines.append("aos  AP:" + ", ".join(f"{value:.2f}" for value in ap_aos[j, :6, i]))
    result = "".join(print_str(line) for line in lines)

    ret_dict = {"result": per_class}
    for tag, table in (("3d", ap_3d), ("bev", ap_bev), ("image", ap_bbox)):
        for level, name in enumerate(("easy", "moderate", "hard")):
            ret_dict['Car_%s_%s' % (tag, name)] = table[0, level, 0]
    return result, ret_dict
