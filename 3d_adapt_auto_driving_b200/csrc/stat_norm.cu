// stat_norm.cu -- Statistical Normalization point rescale for whole scenes on the GPU, sm_100a.
//
// Replaces, for the default options of stat_norm/norm.py:convert (avoid_conflict = align_front = False), the numpy
// chain of rescale_ptc + format_lidar_data per scene (norm.py:186-244, 42-45; utils/kitti_util.py:141-160):
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
                                                               int32_t *__restrict__ box_counts, long long cap, long long cap_out) {
    __shared__ int wcnt[kWarps];
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
                const double s0 = __dmul_rn(p0, bp.scale[0]), s1 = __dmul_rn(p1, bp.scale[1]), s2 = __dmul_rn(p2, bp.scale[2]);
                // np.dot(tmp, R.T) + obj.t : (R.T)[k][j] = R[j][k]
                double q0 = __dmul_rn(s0, bp.R[0]); q0 = __fma_rn(s1, bp.R[1], q0); q0 = __fma_rn(s2, bp.R[2], q0);
                double q1 = __dmul_rn(s0, bp.R[3]); q1 = __fma_rn(s1, bp.R[4], q1); q1 = __fma_rn(s2, bp.R[5], q1);
                double q2 = __dmul_rn(s0, bp.R[6]); q2 = __fma_rn(s1, bp.R[7], q2); q2 = __fma_rn(s2, bp.R[8], q2);
                q0 = __dadd_rn(q0, bp.t[0]); q1 = __dadd_rn(q1, bp.t[1]); q2 = __dadd_rn(q2, bp.t[2]);
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
                                                    reinterpret_cast<float4 *>(out), out_counts, box_counts, cap, cap_out);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
