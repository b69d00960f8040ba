// stat_norm.cu -- Statistical Normalization point rescale for whole scenes on the GPU, sm_100a.
//
// Replaces the numpy chain of rescale_ptc + format_lidar_data per scene (norm.py:186-244, 42-45;
// utils/kitti_util.py:141-160), with all four combinations of convert's avoid_conflict / align_front options:
//   velodyne -> reference -> rectified camera | per Car / Van box: into the box frame, strict in-box mask, scale
//   the in-box points per axis, back to the camera frame | [patch of box 0, patch of box 1, ..., untouched points]
//   -> rectified -> reference -> velodyne | float32 (x, y, z, 1.0) rows of the output .bin.
// The reference computes all of it in float64 through np.dot; every product below is the dgemm order measured on
// numpy (k-sequential FMA chain, first term a plain multiply: fma(a2,b2,fma(a1,b1,a0*b0))) written with explicit
// _rn intrinsics, so the float64 values -- and therefore the float32 rows of the file -- are bit-identical
// (tests/test_stat_norm_gpu.py compares file bytes).  Matrices that numpy builds from scalars (cos / sin of ry,
// inv(R0), the scale factors of get_scale_map) are computed on the host by the same numpy calls and passed in.
//
// One CTA per scene; points are visited in index order and every output list is an ORDERED compaction (ballot +
// shared-memory scan), which is exactly the order boolean-mask indexing gives.  A point inside two boxes appears in
// both patches, as in the reference (each box tests all points).
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;

struct SceneMats {      // 42 doubles per scene
    double v2c_t[12];   // np.transpose(V2C): (4,3) row-major
    double r0[9];       // R0 (3,3)
    double r0_inv[9];   // np.linalg.inv(R0)
    double c2v_t[12];   // np.transpose(C2V): (4,3)
};

struct BoxParams {      // 18 doubles per rescaled box
    double t[3];        // obj.t (float32 promoted)
    double R[9];        // [[c,0,s],[0,1,0],[-s,0,c]] from np.cos / np.sin (ry)
    double half_l, h, half_w;   // obj.l / 2.0, obj.h, obj.w / 2.0
    double scale[3];    // mapping(obj, 1): factors along l, h, w
};

constexpr int kRatios = 11;     // np.arange(1, -0.1, -0.1): the candidates of the avoid_conflict search (norm.py:206)

struct BoxOpts {        // 78 doubles per rescaled box, only for avoid_conflict / align_front (norm.py:205-240)
    double scale[kRatios][3];   // mapping(obj, ratio_r): factors along l, h, w for every candidate ratio
    double shift[kRatios][4];   // align_front: up to two (dx, dz) pairs added to the patch one after the other
    double nshift;              // 0, 1 or 2 pairs (does not depend on the ratio)
};

// sum_k a_k * b[k*ld + j], k-sequential FMA chain, first term a plain multiply (numpy / OpenBLAS dgemm order)
__device__ __forceinline__ double chain3(double a0, double a1, double a2, const double *b, int ld, int j) {
    double t = __dmul_rn(a0, b[j]);
    t = __fma_rn(a1, b[ld + j], t);
    return __fma_rn(a2, b[2 * ld + j], t);
}
__device__ __forceinline__ double chain4(double a0, double a1, double a2, const double *b, int j) {   // a3 = 1 (homogeneous)
    double t = __dmul_rn(a0, b[j]);
    t = __fma_rn(a1, b[3 + j], t);
    t = __fma_rn(a2, b[6 + j], t);
    return __fma_rn(1.0, b[9 + j], t);
}
// np.transpose(np.dot(M, np.transpose(p))): out_i = sum_k M[i][k] * p_k, same chain over k
__device__ __forceinline__ double chain3_rows(const double *M, int i, double p0, double p1, double p2) {
    double t = __dmul_rn(M[i * 3 + 0], p0);
    t = __fma_rn(M[i * 3 + 1], p1, t);
    return __fma_rn(M[i * 3 + 2], p2, t);
}

// rect -> velodyne (project_rect_to_velo) and the float32 row of the .bin
__device__ __forceinline__ float4 rect_to_bin_row(const SceneMats &m, double x, double y, double z) {
    const double r0 = chain3_rows(m.r0_inv, 0, x, y, z), r1 = chain3_rows(m.r0_inv, 1, x, y, z), r2 = chain3_rows(m.r0_inv, 2, x, y, z);
    const double v0 = chain4(r0, r1, r2, m.c2v_t, 0), v1 = chain4(r0, r1, r2, m.c2v_t, 1), v2 = chain4(r0, r1, r2, m.c2v_t, 2);
    return make_float4(__double2float_rn(v0), __double2float_rn(v1), __double2float_rn(v2), 1.0f);
}

__global__ void __launch_bounds__(kThreads) stat_rescale_kernel(const float4 *__restrict__ raw,
                                                               const long long *__restrict__ offsets,
                                                               const SceneMats *__restrict__ mats,
                                                               const BoxParams *__restrict__ boxes,
                                                               const int32_t *__restrict__ box_offsets,
                                                               double *__restrict__ rect, unsigned char *__restrict__ untouched,
                                                               float4 *__restrict__ out, int32_t *__restrict__ out_counts,
                                                               int32_t *__restrict__ box_counts, long long cap, long long cap_out,
                                                               const BoxOpts *__restrict__ opts, int avoid_conflict,
                                                               int32_t *__restrict__ box_ratio) {
    __shared__ int wcnt[kWarps];
    __shared__ double red_lo[kWarps][3], red_hi[kWarps][3];
    __shared__ long long base;
    __shared__ SceneMats m;
    __shared__ BoxParams bp;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < (int)(sizeof(SceneMats) / 8)) reinterpret_cast<double *>(&m)[tid] = reinterpret_cast<const double *>(mats + b)[tid];
    if (tid == 0) base = 0;
    __syncthreads();
    const long long beg = offsets[b], n = offsets[b + 1] - beg;
    raw += beg;
    rect += (size_t)b * cap * 3;
    untouched += (size_t)b * cap;
    out += (size_t)b * cap_out;
    const unsigned lt = (1u << lane) - 1u;
    bool overflow = false;

    // A. velodyne -> rectified camera coordinates (float64), kept for the box passes
    for (long long i = tid; i < n; i += kThreads) {
        const float4 q = __ldg(raw + i);
        const double x = (double)q.x, y = (double)q.y, z = (double)q.z;
        const double f0 = chain4(x, y, z, m.v2c_t, 0), f1 = chain4(x, y, z, m.v2c_t, 1), f2 = chain4(x, y, z, m.v2c_t, 2);
        rect[i * 3 + 0] = chain3_rows(m.r0, 0, f0, f1, f2);
        rect[i * 3 + 1] = chain3_rows(m.r0, 1, f0, f1, f2);
        rect[i * 3 + 2] = chain3_rows(m.r0, 2, f0, f1, f2);
        untouched[i] = 1;
    }
    __syncthreads();

    // B. one ordered pass per rescaled box: the patch of its in-box points, transformed, appended to the output
    for (int bi = box_offsets[b]; bi < box_offsets[b + 1]; ++bi) {
        if (tid < (int)(sizeof(BoxParams) / 8)) reinterpret_cast<double *>(&bp)[tid] = reinterpret_cast<const double *>(boxes + bi)[tid];
        __syncthreads();
        const long long box_base = base;
        // scale factors (and front-alignment shifts) of this box: ratio 1 unless the conflict search backs off
        double sc0 = bp.scale[0], sc1 = bp.scale[1], sc2 = bp.scale[2];
        int ridx = 0;
        if (opts != nullptr && avoid_conflict) {
            // norm.py:205-216.  scaled = local[inside] * mapping(obj, ratio) and its per-axis min / max: the factors are
            // positive and IEEE multiplication is monotone, so min(scaled_k) = fl(min(local_k) * s_k) exactly -- one
            // reduction over the in-box points serves all eleven candidate ratios
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            int n_in = 0, n_clear = 0;
            for (long long i = tid; i < n; i += kThreads) {
                const double d0 = __dsub_rn(rect[i * 3 + 0], bp.t[0]), d1 = __dsub_rn(rect[i * 3 + 1], bp.t[1]),
                             d2 = __dsub_rn(rect[i * 3 + 2], bp.t[2]);
                const double p0 = chain3(d0, d1, d2, bp.R, 3, 0), p1 = chain3(d0, d1, d2, bp.R, 3, 1), p2 = chain3(d0, d1, d2, bp.R, 3, 2);
                const bool foot = p0 > -bp.half_l && p0 < bp.half_l && p2 > -bp.half_w && p2 < bp.half_w && p1 > -bp.h;
                if (foot && p1 < 0.0) {
                    ++n_in;
                    lo[0] = fmin(lo[0], p0); hi[0] = fmax(hi[0], p0);
                    lo[1] = fmin(lo[1], p1); hi[1] = fmax(hi[1], p1);
                    lo[2] = fmin(lo[2], p2); hi[2] = fmax(hi[2], p2);
                }
                n_clear += (foot && p1 < -0.5) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                    hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
                }
                n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
                n_clear += __shfl_xor_sync(0xffffffffu, n_clear, o);
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { red_lo[warp][k] = lo[k]; red_hi[warp][k] = hi[k]; }
                wcnt[warp] = n_in;
            }
            __syncthreads();
            int tot_in = 0;
            for (int w = 0; w < kWarps; ++w) {
                tot_in += wcnt[w];
#pragma unroll
                for (int k = 0; k < 3; ++k) { lo[k] = fmin(lo[k], red_lo[w][k]); hi[k] = fmax(hi[k], red_hi[w][k]); }
            }
            __syncthreads();
            if (lane == 0) wcnt[warp] = n_clear;
            __syncthreads();
            int tot_clear = 0;
            for (int w = 0; w < kWarps; ++w) tot_clear += wcnt[w];
            __syncthreads();
            if (tot_in > 0) {
                const BoxOpts &bo = opts[bi];
                ridx = kRatios - 1;
                for (int r = 0; r < kRatios; ++r) {
                    // bounds of the scaled patch; the lowest 0.5 m are ignored (norm.py:211-213)
                    const double x_lo = __dmul_rn(lo[0], bo.scale[r][0]), x_hi = __dmul_rn(hi[0], bo.scale[r][0]);
                    const double y_lo = __dmul_rn(lo[1], bo.scale[r][1]);
                    const double z_lo = __dmul_rn(lo[2], bo.scale[r][2]), z_hi = __dmul_rn(hi[2], bo.scale[r][2]);
                    int c = 0;
                    for (long long i = tid; i < n; i += kThreads) {
                        const double d0 = __dsub_rn(rect[i * 3 + 0], bp.t[0]), d1 = __dsub_rn(rect[i * 3 + 1], bp.t[1]),
                                     d2 = __dsub_rn(rect[i * 3 + 2], bp.t[2]);
                        const double p0 = chain3(d0, d1, d2, bp.R, 3, 0), p1 = chain3(d0, d1, d2, bp.R, 3, 1),
                                     p2 = chain3(d0, d1, d2, bp.R, 3, 2);
                        c += (p0 > x_lo && p0 < x_hi && p1 > y_lo && p1 < -0.5 && p2 > z_lo && p2 < z_hi) ? 1 : 0;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    if (lane == 0) wcnt[warp] = c;
                    __syncthreads();
                    int swallowed = 0;
                    for (int w = 0; w < kWarps; ++w) swallowed += wcnt[w];
                    __syncthreads();
                    if (swallowed - tot_clear < 10) { ridx = r; break; }      // uniform: every thread sees the same sums
                }
                sc0 = bo.scale[ridx][0]; sc1 = bo.scale[ridx][1]; sc2 = bo.scale[ridx][2];
            }
            if (tid == 0 && box_ratio) box_ratio[bi] = ridx;
        } else if (tid == 0 && box_ratio) {
            box_ratio[bi] = 0;
        }
        double sh[4] = {0.0, 0.0, 0.0, 0.0};
        int nshift = 0;
        if (opts != nullptr) {
            nshift = (int)opts[bi].nshift;
#pragma unroll
            for (int k = 0; k < 4; ++k) sh[k] = opts[bi].shift[ridx][k];
        }
        for (long long i0 = 0; i0 < n; i0 += kThreads) {
            const long long i = i0 + tid;
            bool in = false;
            double p0 = 0, p1 = 0, p2 = 0;
            if (i < n) {
                const double d0 = __dsub_rn(rect[i * 3 + 0], bp.t[0]), d1 = __dsub_rn(rect[i * 3 + 1], bp.t[1]),
                             d2 = __dsub_rn(rect[i * 3 + 2], bp.t[2]);
                p0 = chain3(d0, d1, d2, bp.R, 3, 0);          // np.dot(ptc - obj.t, R)
                p1 = chain3(d0, d1, d2, bp.R, 3, 1);
                p2 = chain3(d0, d1, d2, bp.R, 3, 2);
                in = p0 > -bp.half_l && p0 < bp.half_l && p2 > -bp.half_w && p2 < bp.half_w && p1 > -bp.h && p1 < 0.0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (lane == 0) wcnt[warp] = __popc(bal);
            __syncthreads();
            long long pos = base;
            int total = 0;
            for (int w = 0; w < kWarps; ++w) {
                if (w < warp) pos += wcnt[w];
                total += wcnt[w];
            }
            if (in) {
                pos += __popc(bal & lt);
                const double s0 = __dmul_rn(p0, sc0), s1 = __dmul_rn(p1, sc1), s2 = __dmul_rn(p2, sc2);
                // np.dot(tmp, R.T) + obj.t : (R.T)[k][j] = R[j][k]
                double q0 = __dmul_rn(s0, bp.R[0]); q0 = __fma_rn(s1, bp.R[1], q0); q0 = __fma_rn(s2, bp.R[2], q0);
                double q1 = __dmul_rn(s0, bp.R[3]); q1 = __fma_rn(s1, bp.R[4], q1); q1 = __fma_rn(s2, bp.R[5], q1);
                double q2 = __dmul_rn(s0, bp.R[6]); q2 = __fma_rn(s1, bp.R[7], q2); q2 = __fma_rn(s2, bp.R[8], q2);
                q0 = __dadd_rn(q0, bp.t[0]); q1 = __dadd_rn(q1, bp.t[1]); q2 = __dadd_rn(q2, bp.t[2]);
                // align_front (norm.py:219-240): patch[:, 0] += shift * cos(angle); patch[:, 2] += shift * sin(angle), per rule
                if (nshift > 0) { q0 = __dadd_rn(q0, sh[0]); q2 = __dadd_rn(q2, sh[1]); }
                if (nshift > 1) { q0 = __dadd_rn(q0, sh[2]); q2 = __dadd_rn(q2, sh[3]); }
                if (pos < cap_out) out[pos] = rect_to_bin_row(m, q0, q1, q2);
                else overflow = true;
                untouched[i] = 0;
            }
            __syncthreads();
            if (tid == 0) base += total;
            __syncthreads();
        }
        if (tid == 0) box_counts[bi] = (int32_t)(base - box_base);
        __syncthreads();
    }

    // C. the untouched points, in order
    for (long long i0 = 0; i0 < n; i0 += kThreads) {
        const long long i = i0 + tid;
        const bool keep = i < n && untouched[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        long long pos = base;
        int total = 0;
        for (int w = 0; w < kWarps; ++w) {
            if (w < warp) pos += wcnt[w];
            total += wcnt[w];
        }
        if (keep) {
            pos += __popc(bal & lt);
            if (pos < cap_out) out[pos] = rect_to_bin_row(m, rect[i * 3 + 0], rect[i * 3 + 1], rect[i * 3 + 2]);
            else overflow = true;
        }
        __syncthreads();
        if (tid == 0) base += total;
        __syncthreads();
    }
    const int any_overflow = __syncthreads_or(overflow ? 1 : 0);
    if (tid == 0) out_counts[b] = any_overflow ? -1 : (int32_t)base;
}

}  // namespace

// raw (total, 4) f32 clouds of the batch concatenated, offsets (b + 1) int64; mats (b, 42) f64; boxes (nboxes, 18) f64 with
// box_offsets (b + 1) int32 (only the boxes of the rescaled classes, in label order); rect (b, cap, 3) f64 and untouched
// (b, cap) u8 scratch; out (b, cap_out, 4) f32 = rows of the output .bin; out_counts (b) int32 rows written (-1: cap_out
// too small, retry larger); box_counts (nboxes) int32 in-box points per box (ratio = 1 iff > 0, norm.py:205-216).
PN2_API int pn2_stat_rescale_f64(const float *raw, const long long *offsets, const double *mats, const double *boxes,
                                 const int32_t *box_offsets, double *rect, unsigned char *untouched, float *out,
                                 int32_t *out_counts, int32_t *box_counts, int b, long long cap, long long cap_out,
                                 cudaStream_t stream) {
    if (b < 0 || cap < 0 || cap_out < 0 || (b > 0 && (!raw || !offsets || !mats || !box_offsets || !rect || !untouched || !out || !out_counts)) ||
        (reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        pn2_set_last_error("pn2_stat_rescale_f64: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    stat_rescale_kernel<<<b, kThreads, 0, stream>>>(reinterpret_cast<const float4 *>(raw), offsets,
                                                    reinterpret_cast<const SceneMats *>(mats),
                                                    reinterpret_cast<const BoxParams *>(boxes), box_offsets, rect, untouched,
                                                    reinterpret_cast<float4 *>(out), out_counts, box_counts, cap, cap_out,
                                                    nullptr, 0, nullptr);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// pn2_stat_rescale_f64 with the two options of stat_norm/norm.py:convert (norm.py:205-240):
//   avoid_conflict  per box, the largest ratio of np.arange(1, -0.1, -0.1) whose scaled patch swallows fewer than ten
//                   foreign points (the search of norm.py:205-216, on the device: one min / max reduction + one count per
//                   candidate); box_ratio (nboxes) int32 receives the index of the chosen ratio (0 = ratio 1)
//   align_front     the patch is shifted so that the face nearest to the sensor stays in place (norm.py:219-240)
// box_opts (nboxes, 78) f64 per rescaled box: scale[11][3] = mapping(obj, ratio_r) for the eleven candidate ratios,
// shift[11][4] = up to two (dx, dz) pairs per ratio (zeros without align_front), nshift = how many pairs apply -- all
// computed on the host by the reference's own numpy expressions, so the kernel only reproduces float64 products and sums.
PN2_API int pn2_stat_rescale_opts_f64(const float *raw, const long long *offsets, const double *mats, const double *boxes,
                                      const int32_t *box_offsets, const double *box_opts, int avoid_conflict, double *rect,
                                      unsigned char *untouched, float *out, int32_t *out_counts, int32_t *box_counts,
                                      int32_t *box_ratio, int b, long long cap, long long cap_out, cudaStream_t stream) {
    if (b < 0 || cap < 0 || cap_out < 0 || (b > 0 && (!raw || !offsets || !mats || !box_offsets || !rect || !untouched || !out || !out_counts || !box_opts)) ||
        (reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        pn2_set_last_error("pn2_stat_rescale_opts_f64: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    stat_rescale_kernel<<<b, kThreads, 0, stream>>>(reinterpret_cast<const float4 *>(raw), offsets,
                                                    reinterpret_cast<const SceneMats *>(mats),
                                                    reinterpret_cast<const BoxParams *>(boxes), box_offsets, rect, untouched,
                                                    reinterpret_cast<float4 *>(out), out_counts, box_counts, cap, cap_out,
                                                    reinterpret_cast<const BoxOpts *>(box_opts), avoid_conflict, box_ratio);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
