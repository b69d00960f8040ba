// rotate_iou.cu -- rotated-rectangle IoU matrix for KITTI AP, sm_100a.
//
// Replaces evaluate/rotate_iou.py:16-291 (numba.cuda kernel rotate_iou_kernel_eval and its device
// functions) behind pn2_rotate_iou_eval_f32.  boxes (N,5), query_boxes (K,5) rows [cx, cy, w, h,
// angle] -> out (N,K); criterion -1: IoU, 0: inter / area(query), 1: inter / area(box), 2: inter.
//
// The reference's bits are defined by numba's type inference plus what NVVM and ptxas fuse, so
// the arithmetic below is pinned with _rn intrinsics to the compiled reference (PTX dumped by
// numba 0.65 on a B200, tests/golden/rotate_iou_numba_sm90.ptx, and the SASS ptxas 12.9 makes of it):
//   * everything is float32 except: the polygon centroid (sum / count as a float64 division), the
//     triangle areas (each f32 cross product is halved, abs'ed and accumulated in float64) and the
//     final ratio (float64 division, then rounded to float32);
//   * a*b + c*d and a*b - c*d run as fma(a, b, +-rn(c*d)) with the FIRST product fused, except the
//     two mixed dot products of point_in_quadrilateral, where NVVM fused the second one;
//   * products that are only compared (the orientation tests of line_segment_intersection) are
//     rounded separately;  sqrt and the vertex normalisation are IEEE (sqrt.rn, div.rn).
// Layout: one thread per (box, query) pair, 64 x 4 pairs per CTA with both box sets staged in
// shared memory; the query index is the fast thread index so result rows are written coalesced.
#include "common.cuh"
#include <math.h>

namespace {

// rotate_iou.py:203-227 rbbox_to_corners
__device__ __forceinline__ void box_corners(const float *rb, float *c) {
    const float a_cos = cosf(rb[4]), a_sin = sinf(rb[4]);
    const float cx = rb[0], cy = rb[1];
    const float xh = __fmul_rn(rb[2], 0.5f), yh = __fmul_rn(rb[3], 0.5f);    // -x_d / 2 etc. are exact
    const float px[4] = {-xh, -xh, xh, xh};
    const float py[4] = {-yh, yh, yh, -yh};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c[2 * i] = __fadd_rn(__fmaf_rn(px[i], a_cos, __fmul_rn(a_sin, py[i])), cx);
        c[2 * i + 1] = __fadd_rn(__fmaf_rn(py[i], a_cos, -__fmul_rn(px[i], a_sin)), cy);
    }
}

// rotate_iou.py:160-176
__device__ __forceinline__ bool point_in_quad(float pt_x, float pt_y, const float *q) {
    const float ab0 = __fsub_rn(q[2], q[0]), ab1 = __fsub_rn(q[3], q[1]);
    const float ad0 = __fsub_rn(q[6], q[0]), ad1 = __fsub_rn(q[7], q[1]);
    const float ap0 = __fsub_rn(pt_x, q[0]), ap1 = __fsub_rn(pt_y, q[1]);
    const float abab = __fmaf_rn(ab0, ab0, __fmul_rn(ab1, ab1));
    const float abap = __fmaf_rn(ab1, ap1, __fmul_rn(ab0, ap0));
    const float adad = __fmaf_rn(ad0, ad0, __fmul_rn(ad1, ad1));
    const float adap = __fmaf_rn(ad1, ap1, __fmul_rn(ad0, ap0));
    return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}

// rotate_iou.py:72-115: edge i of pts1 against edge j of pts2
__device__ __forceinline__ bool segment_intersection(const float *pts1, const float *pts2, int i, int j, float *out) {
    const float A0 = pts1[2 * i], A1 = pts1[2 * i + 1];
    const float B0 = pts1[2 * ((i + 1) & 3)], B1 = pts1[2 * ((i + 1) & 3) + 1];
    const float C0 = pts2[2 * j], C1 = pts2[2 * j + 1];
    const float D0 = pts2[2 * ((j + 1) & 3)], D1 = pts2[2 * ((j + 1) & 3) + 1];
    const float BA0 = __fsub_rn(B0, A0), BA1 = __fsub_rn(B1, A1);
    const float DA0 = __fsub_rn(D0, A0), CA0 = __fsub_rn(C0, A0);
    const float DA1 = __fsub_rn(D1, A1), CA1 = __fsub_rn(C1, A1);
    const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
    const bool bcd = __fmul_rn(__fsub_rn(D1, B1), __fsub_rn(C0, B0)) > __fmul_rn(__fsub_rn(C1, B1), __fsub_rn(D0, B0));
    if (acd == bcd) return false;
    const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
    const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
    if (abc == abd) return false;
    const float DC0 = __fsub_rn(D0, C0), DC1 = __fsub_rn(D1, C1);
    const float ABBA = __fmaf_rn(A0, B1, -__fmul_rn(B0, A1));
    const float CDDC = __fmaf_rn(C0, D1, -__fmul_rn(D0, C1));
    const float DH = __fmaf_rn(BA1, DC0, -__fmul_rn(BA0, DC1));
    const float Dx = __fmaf_rn(ABBA, DC0, -__fmul_rn(BA0, CDDC));
    const float Dy = __fmaf_rn(ABBA, DC1, -__fmul_rn(BA1, CDDC));
    out[0] = __fdiv_rn(Dx, DH);
    out[1] = __fdiv_rn(Dy, DH);
    return true;
}

constexpr int kMaxPts = 24;   // 8 corners + 16 edge crossings (the reference's local array holds 8 points;
                              // more than 8 is out-of-bounds there, i.e. undefined)

// rotate_iou.py:230-244 inter(): corners, intersection polygon, vertex sort, fan triangulation
__device__ double inter_area(const float *rbox1, const float *rbox2) {
    float c1[8], c2[8], pts[2 * kMaxPts];
    box_corners(rbox1, c1);
    box_corners(rbox2, c2);
    int n = 0;
    for (int i = 0; i < 4; ++i) {                                          // :179-189
        if (point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) { pts[2 * n] = c1[2 * i]; pts[2 * n + 1] = c1[2 * i + 1]; ++n; }
        if (point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) { pts[2 * n] = c2[2 * i]; pts[2 * n + 1] = c2[2 * i + 1]; ++n; }
    }
    for (int i = 0; i < 4; ++i)                                            // :190-198
        for (int j = 0; j < 4; ++j) {
            float t[2];
            if (segment_intersection(c1, c2, i, j, t)) { pts[2 * n] = t[0]; pts[2 * n + 1] = t[1]; ++n; }
        }
    if (n > 0) {                                                           // :32-69 sort_vertex_in_convex_polygon
        float s0 = 0.f, s1 = 0.f;
        for (int i = 0; i < n; ++i) { s0 = __fadd_rn(s0, pts[2 * i]); s1 = __fadd_rn(s1, pts[2 * i + 1]); }
        const float m0 = (float)__ddiv_rn((double)s0, (double)n);
        const float m1 = (float)__ddiv_rn((double)s1, (double)n);
        float vs[kMaxPts];
        for (int i = 0; i < n; ++i) {
            float v0 = __fsub_rn(pts[2 * i], m0), v1 = __fsub_rn(pts[2 * i + 1], m1);
            const float d = __fsqrt_rn(__fmaf_rn(v0, v0, __fmul_rn(v1, v1)));
            v0 = __fdiv_rn(v0, d);
            v1 = __fdiv_rn(v1, d);
            if (v1 < 0.f) v0 = __fsub_rn(-2.f, v0);
            vs[i] = v0;
        }
        for (int i = 1; i < n; ++i) {                                      // insertion sort by vs
            if (vs[i - 1] > vs[i]) {
                const float temp = vs[i], tx = pts[2 * i], ty = pts[2 * i + 1];
                int j = i;
                while (j > 0 && vs[j - 1] > temp) {
                    vs[j] = vs[j - 1];
                    pts[2 * j] = pts[2 * j - 2];
                    pts[2 * j + 1] = pts[2 * j - 1];
                    --j;
                }
                vs[j] = temp;
                pts[2 * j] = tx;
                pts[2 * j + 1] = ty;
            }
        }
    }
    double area_val = 0.0;                                                 // :16-29 area / trangle_area
    for (int i = 0; i < n - 2; ++i) {
        const float *a = pts, *b = pts + 2 * i + 2, *c = pts + 2 * i + 4;
        const float cr = __fmaf_rn(__fsub_rn(a[0], c[0]), __fsub_rn(b[1], c[1]),
                                   -__fmul_rn(__fsub_rn(a[1], c[1]), __fsub_rn(b[0], c[0])));
        area_val = __dadd_rn(area_val, fabs(__dmul_rn((double)cr, 0.5)));
    }
    return area_val;
}

// rotate_iou.py:247-259 devRotateIoUEval(rbox1 = query box, rbox2 = box)
__device__ __forceinline__ float rotate_iou_eval(const float *rbox1, const float *rbox2, int criterion) {
    const float area1 = __fmul_rn(rbox1[2], rbox1[3]);
    const float area2 = __fmul_rn(rbox2[2], rbox2[3]);
    const double ai = inter_area(rbox1, rbox2);
    double r;
    if (criterion == -1) r = __ddiv_rn(ai, __dsub_rn((double)__fadd_rn(area1, area2), ai));
    else if (criterion == 0) r = __ddiv_rn(ai, (double)area1);
    else if (criterion == 1) r = __ddiv_rn(ai, (double)area2);
    else r = ai;
    return (float)r;
}

constexpr int kQ = 64, kB = 4;   // queries x boxes per CTA

__global__ void __launch_bounds__(kQ *kB) rotate_iou_kernel(int n, int k, const float *__restrict__ boxes,
                                                            const float *__restrict__ qboxes, float *__restrict__ out,
                                                            int criterion) {
    __shared__ float sq[kQ][5], sb[kB][5];
    const int tq = threadIdx.x, tb = threadIdx.y;
    const int t = tb * kQ + tq;
    for (int e = t; e < kQ * 5; e += kQ * kB) {
        const int gq = blockIdx.x * kQ + e / 5;
        sq[e / 5][e % 5] = gq < k ? qboxes[(size_t)gq * 5 + e % 5] : 0.f;
    }
    if (t < kB * 5) {
        const int gb = blockIdx.y * kB + t / 5;
        sb[t / 5][t % 5] = gb < n ? boxes[(size_t)gb * 5 + t % 5] : 0.f;
    }
    __syncthreads();
    const int iq = blockIdx.x * kQ + tq, ib = blockIdx.y * kB + tb;
    if (iq >= k || ib >= n) return;
    out[(size_t)ib * k + iq] = rotate_iou_eval(sq[tq], sb[tb], criterion);
}

}  // namespace

// evaluate/rotate_iou.py:294-329 rotate_iou_gpu_eval, device part.  boxes (n,5), qboxes (k,5) f32 ->
// out (n,k) f32.
PN2_API int pn2_rotate_iou_eval_f32(const float *boxes, int n, const float *qboxes, int k, float *out, int criterion,
                                    cudaStream_t stream) {
    if (n < 0 || k < 0 || (n > 0 && k > 0 && (!boxes || !qboxes || !out))) {
        pn2_set_last_error("pn2_rotate_iou_eval_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (n == 0 || k == 0) return PN2_OK;
    dim3 grid(pn2_divup(k, kQ), pn2_divup(n, kB)), block(kQ, kB);
    if (grid.y > 65535) {
        pn2_set_last_error("pn2_rotate_iou_eval_f32: more than 262140 boxes");
        return PN2_ERR_UNSUPPORTED;
    }
    rotate_iou_kernel<<<grid, block, 0, stream>>>(n, k, boxes, qboxes, out, criterion);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
