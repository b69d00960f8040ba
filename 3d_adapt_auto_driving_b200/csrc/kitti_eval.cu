// kitti_eval.cu -- the KITTI AP evaluator's inner loops, the consumer of rotate_iou.
//
// Replaces the numba kernels of evaluate/eval2.py:
//   image_box_overlap        :104-127   (host, float64)                      -> pn2_eval_image_box_overlap
//   d3_box_overlap_kernel    :136-162   (numba parallel CPU loop)            -> pn2_d3_overlap_f64  (DEVICE kernel)
//   compute_statistics_jit   :172-298   (greedy gt <-> detection matching)   -> compute_statistics() below
//   fused_compute_statistics :311-358   (per image x per threshold driver)   -> pn2_eval_fused_statistics
// and the per-image threshold collection loop of eval_class (:506-520)       -> pn2_eval_collect_thresholds.
// The matching is a sequential greedy pass per (image, threshold): host code, as in the reference, but native and
// driven over whole parts of the dataset per call; the 3-D overlap (height overlap x rotated BEV intersection from
// pn2_rotate_iou_f32, criterion 2) is one thread per pair on the device, in float64 like the numba loop.
#include "common.cuh"
#include <math.h>
#include <stdint.h>
#include <vector>

namespace {

// eval2.py:104-127 (criterion -1: IoU, 0: / area(box), 1: / area(query), else 1.0); overlaps (N, K) row-major
void image_box_overlap(const double *boxes, long long n, const double *qboxes, long long k, int criterion, double *out) {
    for (long long i = 0; i < n * k; ++i) out[i] = 0.0;
    for (long long kk = 0; kk < k; ++kk) {
        const double *q = qboxes + kk * 4;
        const double qarea = (q[2] - q[0]) * (q[3] - q[1]);
        for (long long nn = 0; nn < n; ++nn) {
            const double *b = boxes + nn * 4;
            const double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]);
            if (iw > 0) {
                const double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]);
                if (ih > 0) {
                    double ua;
                    if (criterion == -1) ua = (b[2] - b[0]) * (b[3] - b[1]) + qarea - iw * ih;
                    else if (criterion == 0) ua = (b[2] - b[0]) * (b[3] - b[1]);
                    else if (criterion == 1) ua = qarea;
                    else ua = 1.0;
                    out[nn * k + kk] = iw * ih / ua;
                }
            }
        }
    }
}

struct Stats {
    long long tp, fp, fn;
    double similarity;     // -1 = undefined (no tp and no fp), eval2.py:295-296
};

// eval2.py:172-298.  overlaps: (det, gt) view with row stride ldo; gt_datas (gt, 5) [bbox4, alpha]; dt_datas (det, 6)
// [bbox4, alpha, score]; thresholds_out (optional, gt_size entries) receives the scores of the true positives.
Stats compute_statistics(const double *overlaps, long long ldo, const double *gt_datas, long long gt_size,
                         const double *dt_datas, long long det_size, const long long *ignored_gt,
                         const long long *ignored_det, const double *dc_bboxes, long long n_dc, int metric,
                         double min_overlap, double thresh, bool compute_fp, bool compute_aos, double *thresholds_out,
                         long long *n_thresholds, std::vector<char> &assigned, std::vector<char> &ign_thr,
                         std::vector<double> &delta, std::vector<double> &work) {
    const double NO_DETECTION = -10000000.0;
    assigned.assign((size_t)det_size, 0);
    ign_thr.assign((size_t)det_size, 0);
    if (compute_fp)
        for (long long i = 0; i < det_size; ++i)
            if (dt_datas[i * 6 + 5] < thresh) ign_thr[(size_t)i] = 1;
    long long tp = 0, fp = 0, fn = 0, thresh_idx = 0, delta_idx = 0;
    double similarity = 0.0;
    delta.assign((size_t)gt_size, 0.0);
    for (long long i = 0; i < gt_size; ++i) {
        if (ignored_gt[i] == -1) continue;
        long long det_idx = -1;
        double valid_detection = NO_DETECTION, max_overlap = 0.0;
        bool assigned_ignored_det = false;
        for (long long j = 0; j < det_size; ++j) {
            if (ignored_det[j] == -1) continue;
            if (assigned[(size_t)j]) continue;
            if (ign_thr[(size_t)j]) continue;
            const double overlap = overlaps[j * ldo + i];
            const double dt_score = dt_datas[j * 6 + 5];
            if (!compute_fp && overlap > min_overlap && dt_score > valid_detection) {
                det_idx = j;
                valid_detection = dt_score;
            } else if (compute_fp && overlap > min_overlap && (overlap > max_overlap || assigned_ignored_det) &&
                       ignored_det[j] == 0) {
                max_overlap = overlap;
                det_idx = j;
                valid_detection = 1;
                assigned_ignored_det = false;
            } else if (compute_fp && overlap > min_overlap && valid_detection == NO_DETECTION && ignored_det[j] == 1) {
                det_idx = j;
                valid_detection = 1;
                assigned_ignored_det = true;
            }
        }
        if (valid_detection == NO_DETECTION && ignored_gt[i] == 0) {
            fn += 1;
        } else if (valid_detection != NO_DETECTION && (ignored_gt[i] == 1 || ignored_det[det_idx] == 1)) {
            assigned[(size_t)det_idx] = 1;
        } else if (valid_detection != NO_DETECTION) {
            tp += 1;
            if (thresholds_out) thresholds_out[thresh_idx] = dt_datas[det_idx * 6 + 5];
            thresh_idx += 1;
            if (compute_aos) {
                delta[(size_t)delta_idx] = gt_datas[i * 5 + 4] - dt_datas[det_idx * 6 + 4];
                delta_idx += 1;
            }
            assigned[(size_t)det_idx] = 1;
        }
    }
    if (compute_fp) {
        for (long long i = 0; i < det_size; ++i)
            if (!(assigned[(size_t)i] || ignored_det[i] == -1 || ignored_det[i] == 1 || ign_thr[(size_t)i])) fp += 1;
        long long nstuff = 0;
        if (metric == 0 && n_dc > 0 && det_size > 0) {
            // overlaps of the detections with the DontCare boxes, criterion 0 (dt_bboxes = dt_datas[:, :4])
            work.assign((size_t)(det_size * 4 + det_size * n_dc), 0.0);
            double *dtb = work.data(), *ov = work.data() + det_size * 4;
            for (long long i = 0; i < det_size; ++i)
                for (int c = 0; c < 4; ++c) dtb[i * 4 + c] = dt_datas[i * 6 + c];
            image_box_overlap(dtb, det_size, dc_bboxes, n_dc, 0, ov);
            for (long long i = 0; i < n_dc; ++i) {
                for (long long j = 0; j < det_size; ++j) {
                    if (assigned[(size_t)j]) continue;
                    if (ignored_det[j] == -1 || ignored_det[j] == 1) continue;
                    if (ign_thr[(size_t)j]) continue;
                    if (ov[j * n_dc + i] > min_overlap) {
                        assigned[(size_t)j] = 1;
                        nstuff += 1;
                    }
                }
            }
        }
        fp -= nstuff;
        if (compute_aos) {
            if (tp > 0 || fp > 0) {
                double sum = 0.0;                       // np.sum over [0]*fp + [(1 + cos(delta)) / 2 ...], in order
                for (long long i = 0; i < delta_idx; ++i) sum += (1.0 + cos(delta[(size_t)i])) / 2.0;
                similarity = sum;
            } else {
                similarity = -1.0;
            }
        }
    }
    if (n_thresholds) *n_thresholds = thresh_idx;
    Stats s = {tp, fp, fn, similarity};
    return s;
}

// eval2.py:136-162 on the device: boxes / qboxes (·, 7) float64 camera boxes [x, y, z, l, h, w, ry], rinc (N, K) the
// rotated BEV intersection areas (float32 from pn2_rotate_iou_f32, criterion 2) -> out (N, K) float64.
__global__ void d3_overlap_kernel(const double *__restrict__ boxes, long long n, const double *__restrict__ qboxes,
                                  long long k, const float *__restrict__ rinc, int criterion, double *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const long long i = t / k, j = t - i * k;
    const double r = (double)rinc[t];
    double o = r;                                        // numba leaves rinc untouched where it is not > 0
    if (r > 0) {
        const double *b = boxes + i * 7, *q = qboxes + j * 7;
        const double iw = fmin(b[1], q[1]) - fmax(b[1] - b[4], q[1] - q[4]);
        if (iw > 0) {
            const double area1 = b[3] * b[4] * b[5];
            const double area2 = q[3] * q[4] * q[5];
            const double inc = iw * r;
            double ua;
            if (criterion == -1) ua = area1 + area2 - inc;
            else if (criterion == 0) ua = area1;
            else if (criterion == 1) ua = area2;
            else ua = inc;
            o = inc / ua;
        } else {
            o = 0.0;
        }
    }
    out[t] = o;
}

}  // namespace

PN2_API int pn2_eval_image_box_overlap(const double *boxes, long long n, const double *qboxes, long long k, int criterion,
                                       double *out) {
    if (n < 0 || k < 0 || (n * k > 0 && (!boxes || !qboxes || !out))) {
        pn2_set_last_error("pn2_eval_image_box_overlap: bad argument");
        return PN2_ERR_INVALID;
    }
    image_box_overlap(boxes, n, qboxes, k, criterion, out);
    return PN2_OK;
}

PN2_API int pn2_d3_overlap_f64(const double *boxes, long long n, const double *qboxes, long long k, const float *rinc,
                               int criterion, double *out, cudaStream_t stream) {
    if (n < 0 || k < 0 || (n * k > 0 && (!boxes || !qboxes || !rinc || !out))) {
        pn2_set_last_error("pn2_d3_overlap_f64: bad argument");
        return PN2_ERR_INVALID;
    }
    if (n * k == 0) return PN2_OK;
    const long long total = n * k;
    d3_overlap_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(boxes, n, qboxes, k, rinc, criterion, out);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// One part of the dataset (eval2.py:522-550): overlaps (total_dt, total_gt) of the part, images described by
// gt_nums / dt_nums / dc_nums (n_img each) over the concatenated gt_datas (·,5), dt_datas (·,6), dontcares (·,4),
// ignored_gts, ignored_dets.  pr (n_thresholds, 4) += [tp, fp, fn, similarity] per threshold.
PN2_API int pn2_eval_fused_statistics(const double *overlaps, long long total_dt, long long total_gt, double *pr,
                                      const long long *gt_nums, const long long *dt_nums, const long long *dc_nums,
                                      long long n_img, const double *gt_datas, const double *dt_datas,
                                      const double *dontcares, const long long *ignored_gts,
                                      const long long *ignored_dets, int metric, double min_overlap,
                                      const double *thresholds, long long n_thresholds, int compute_aos) {
    if (n_img < 0 || n_thresholds < 0 || (n_img > 0 && (!gt_nums || !dt_nums || !dc_nums)) || (n_thresholds > 0 && !pr)) {
        pn2_set_last_error("pn2_eval_fused_statistics: bad argument");
        return PN2_ERR_INVALID;
    }
    std::vector<char> assigned, ign_thr;
    std::vector<double> delta, work;
    long long gt_num = 0, dt_num = 0, dc_num = 0;
    for (long long i = 0; i < n_img; ++i) {
        if (gt_num + gt_nums[i] > total_gt || dt_num + dt_nums[i] > total_dt) {
            pn2_set_last_error("pn2_eval_fused_statistics: image extents exceed the overlap matrix");
            return PN2_ERR_INVALID;
        }
        for (long long t = 0; t < n_thresholds; ++t) {
            const Stats s = compute_statistics(overlaps + dt_num * total_gt + gt_num, total_gt, gt_datas + gt_num * 5,
                                               gt_nums[i], dt_datas + dt_num * 6, dt_nums[i], ignored_gts + gt_num,
                                               ignored_dets + dt_num, dontcares + dc_num * 4, dc_nums[i], metric,
                                               min_overlap, thresholds[t], true, compute_aos != 0, nullptr, nullptr,
                                               assigned, ign_thr, delta, work);
            pr[t * 4 + 0] += (double)s.tp;
            pr[t * 4 + 1] += (double)s.fp;
            pr[t * 4 + 2] += (double)s.fn;
            if (s.similarity != -1.0) pr[t * 4 + 3] += s.similarity;
        }
        gt_num += gt_nums[i];
        dt_num += dt_nums[i];
        dc_num += dc_nums[i];
    }
    return PN2_OK;
}

// The first pass of eval_class (eval2.py:506-520) over one part: compute_statistics(thresh = 0, compute_fp = False)
// per image; the scores of the true positives are appended to thresholds_out (capacity total_gt), *n_out = count.
PN2_API int pn2_eval_collect_thresholds(const double *overlaps, long long total_dt, long long total_gt,
                                        const long long *gt_nums, const long long *dt_nums, const long long *dc_nums,
                                        long long n_img, const double *gt_datas, const double *dt_datas,
                                        const double *dontcares, const long long *ignored_gts,
                                        const long long *ignored_dets, int metric, double min_overlap,
                                        double *thresholds_out, long long *n_out) {
    if (n_img < 0 || !n_out || (n_img > 0 && (!gt_nums || !dt_nums || !dc_nums))) {
        pn2_set_last_error("pn2_eval_collect_thresholds: bad argument");
        return PN2_ERR_INVALID;
    }
    std::vector<char> assigned, ign_thr;
    std::vector<double> delta, work;
    long long gt_num = 0, dt_num = 0, dc_num = 0, n = 0;
    for (long long i = 0; i < n_img; ++i) {
        if (gt_num + gt_nums[i] > total_gt || dt_num + dt_nums[i] > total_dt) {
            pn2_set_last_error("pn2_eval_collect_thresholds: image extents exceed the overlap matrix");
            return PN2_ERR_INVALID;
        }
        long long got = 0;
        compute_statistics(overlaps + dt_num * total_gt + gt_num, total_gt, gt_datas + gt_num * 5, gt_nums[i],
                           dt_datas + dt_num * 6, dt_nums[i], ignored_gts + gt_num, ignored_dets + dt_num,
                           dontcares + dc_num * 4, dc_nums[i], metric, min_overlap, 0.0, false, false, thresholds_out + n,
                           &got, assigned, ign_thr, delta, work);
        n += got;
        gt_num += gt_nums[i];
        dt_num += dt_nums[i];
        dc_num += dc_nums[i];
    }
    *n_out = n;
    return PN2_OK;
}

// compute_statistics_jit itself (eval2.py:172-298) for ONE image, every mode: out4 = {tp, fp, fn, similarity} with the
// reference's conventions (similarity 0 unless compute_fp && compute_aos, -1 when there is neither a tp nor a fp),
// thresholds_out (capacity gt_size) = scores of the true positives, *n_thresholds = tp.  overlaps is (det, gt) with
// row stride ldo.
PN2_API int pn2_eval_image_statistics(const double *overlaps, long long ldo, const double *gt_datas, long long gt_size,
                                      const double *dt_datas, long long det_size, const long long *ignored_gt,
                                      const long long *ignored_det, const double *dc_bboxes, long long n_dc, int metric,
                                      double min_overlap, double thresh, int compute_fp, int compute_aos, double *out4,
                                      double *thresholds_out, long long *n_thresholds) {
    if (gt_size < 0 || det_size < 0 || n_dc < 0 || !out4 || !n_thresholds || ldo < gt_size ||
        (gt_size > 0 && (!gt_datas || !ignored_gt || !thresholds_out)) || (det_size > 0 && (!dt_datas || !ignored_det)) ||
        (gt_size * det_size > 0 && !overlaps) || (n_dc > 0 && !dc_bboxes)) {
        pn2_set_last_error("pn2_eval_image_statistics: bad argument");
        return PN2_ERR_INVALID;
    }
    std::vector<char> assigned, ign_thr;
    std::vector<double> delta, work;
    const Stats s = compute_statistics(overlaps, ldo, gt_datas, gt_size, dt_datas, det_size, ignored_gt, ignored_det,
                                       dc_bboxes, n_dc, metric, min_overlap, thresh, compute_fp != 0, compute_aos != 0,
                                       thresholds_out, n_thresholds, assigned, ign_thr, delta, work);
    out4[0] = (double)s.tp;
    out4[1] = (double)s.fp;
    out4[2] = (double)s.fn;
    out4[3] = s.similarity;
    return PN2_OK;
}
