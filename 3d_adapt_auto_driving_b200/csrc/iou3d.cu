// iou3d.cu -- rotated / axis-aligned BEV overlap and NMS for sm_100a.
//
// Replaces pointrcnn/lib/utils/iou3d/src/iou3d_kernel.cu:223-348 (boxes_overlap_kernel,
// boxes_iou_bev_kernel, nms_kernel, nms_normal_kernel) and the host greedy pass of
// iou3d.cpp:73-169 behind pn2_boxes_overlap_bev_f32 / pn2_boxes_iou_bev_f32 / pn2_nms_bev_f32.
//
// NMS design.  The reference materialises the full n x n/64 suppression bit matrix
// (n = 6300: 39.7 M IoU evaluations, 5 MB), cudaMalloc's it per call, copies it to the host
// with a blocking cudaMemcpy and runs the greedy pass on the CPU (iou3d.cpp:87-116) -- 32
// times per batch in the proposal layer, which then keeps only the first 70 / 30 survivors
// (proposal_layer.py:112).  Here one CTA per NMS problem runs the greedy pass ON THE DEVICE
// and evaluates a suppression row lazily, only for boxes that are actually kept, and stops
// after `max_keep` survivors: 70 x 6300 IoUs instead of 39.7 M, no scratch, no host round
// trip, data-dependent sizes stay on the device (counts[] in, num[] out) so the whole
// proposal stage can be enqueued without a synchronisation.  The result is identical to the
// mask + greedy formulation: box j is suppressed iff some kept i < j has iou(i, j) > thresh,
// and iou(i, j) is evaluated with the same operand order (row box first).
//
// Overlap arithmetic: same formulas with every rounding pinned by _rn intrinsics to what nvcc
// emits for the reference source, so that areas agree bit for bit (checked against the
// reference kernels on the GPU, tests/test_iou3d_roipool_gpu.py).
#include "common.cuh"
#include <math.h>

namespace {

constexpr float kEps = 1e-8f;

struct P2 {
    float x, y;
};

// Arithmetic pinning.  The bits of an overlap area depend on which products get fused into
// FMAs when the reference source is compiled -- by nvcc's front end AND by ptxas, which fuses
// non-.rn mul/sub pairs of the PTX.  Read from the SASS of the unmodified iou3d_kernel.cu (nvcc 12.9
// -O2, sm_100a; oracle/_ref/libpn2_legacy.so) every "a*b - c*d" and "a*b + c*d" is executed as
//   fma(a, b, -+rn(c*d))       (minuend product fused, the other product rounded first)
// with one exception: s2 = cross(p1,q1,p0) and s5 = cross(q1,p1,p0) of a segment test share
// their two products, which are therefore both rounded and then subtracted.  Centre +- terms
// fused with *0.5 are exact either way.  Everything below is written with _rn intrinsics so
// that no compiler version can contract or re-associate it differently.
__device__ __forceinline__ float diffprod(float a, float b, float c, float d) {   // a*b - c*d as compiled
    return __fmaf_rn(a, b, -__fmul_rn(c, d));
}
__device__ __forceinline__ float cross3(const P2 &p1, const P2 &p2, const P2 &p0) {
    return diffprod(__fsub_rn(p1.x, p0.x), __fsub_rn(p2.y, p0.y), __fsub_rn(p2.x, p0.x), __fsub_rn(p1.y, p0.y));
}

__device__ __forceinline__ bool boxes_apart(const P2 &p1, const P2 &p2, const P2 &q1, const P2 &q2) {
    const bool touch = fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
                       fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
    return !touch;
}

// proper intersection of segment p0-p1 with q0-q1 (iou3d_kernel.cu:66-96)
__device__ __forceinline__ bool seg_intersect(const P2 &p1, const P2 &p0, const P2 &q1, const P2 &q0, P2 &ans) {
    if (boxes_apart(p0, p1, q0, q1)) return false;
    const float s1 = cross3(q0, p1, p0);
    // s2 = (p1.x-p0.x)(q1.y-p0.y) - (q1.x-p0.x)(p1.y-p0.y) and s5 = -s2 share both products
    const float pa = __fmul_rn(__fsub_rn(p1.x, p0.x), __fsub_rn(q1.y, p0.y));
    const float pb = __fmul_rn(__fsub_rn(q1.x, p0.x), __fsub_rn(p1.y, p0.y));
    const float s2 = __fsub_rn(pa, pb);
    const float s3 = cross3(p0, q1, q0);
    const float s4 = cross3(q1, p1, q0);
    if (!(__fmul_rn(s1, s2) > 0 && __fmul_rn(s3, s4) > 0)) return false;
    const float s5 = __fsub_rn(pb, pa);
    const float den = __fsub_rn(s5, s1);
    if (fabsf(den) > kEps) {
        ans.x = __fdiv_rn(diffprod(s5, q0.x, s1, q1.x), den);
        ans.y = __fdiv_rn(diffprod(s5, q0.y, s1, q1.y), den);
    } else {
        const float a0 = __fsub_rn(p0.y, p1.y), b0 = __fsub_rn(p1.x, p0.x);
        const float c0 = diffprod(p0.x, p1.y, p1.x, p0.y);
        const float a1 = __fsub_rn(q0.y, q1.y), b1 = __fsub_rn(q1.x, q0.x);
        const float c1 = diffprod(q0.x, q1.y, q1.x, q0.y);
        const float D = diffprod(a0, b1, a1, b0);
        ans.x = __fdiv_rn(diffprod(b0, c1, b1, c0), D);
        ans.y = __fdiv_rn(diffprod(a1, c0, a0, c1), D);
    }
    return true;
}

// rotation about a centre with the reference's rounding
__device__ __forceinline__ void spin(const P2 &c, float angle_cos, float angle_sin, P2 &p) {
    const float dx = __fsub_rn(p.x, c.x), dy = __fsub_rn(p.y, c.y);
    p.x = __fadd_rn(__fmaf_rn(angle_cos, dx, __fmul_rn(angle_sin, dy)), c.x);
    p.y = __fadd_rn(diffprod(angle_cos, dy, angle_sin, dx), c.y);
}

// point inside the (un-rotated extent of the) box after rotating it back (iou3d_kernel.cu:50-64)
__device__ __forceinline__ bool inside_box(const float *box, const P2 &p) {
    const float margin = 1e-5f;
    P2 c, q = p;
    c.x = __fmul_rn(__fadd_rn(box[0], box[2]), 0.5f);
    c.y = __fmul_rn(__fadd_rn(box[1], box[3]), 0.5f);
    spin(c, cosf(-box[4]), sinf(-box[4]), q);
    return q.x > __fsub_rn(box[0], margin) && q.x < __fadd_rn(box[2], margin) && q.y > __fsub_rn(box[1], margin) &&
           q.y < __fadd_rn(box[3], margin);
}

// intersection area of two rotated rectangles [x1,y1,x2,y2,angle] (iou3d_kernel.cu:108-212)
__device__ float rot_overlap(const float *box_a, const float *box_b) {
    const float ax1 = box_a[0], ay1 = box_a[1], ax2 = box_a[2], ay2 = box_a[3], aang = box_a[4];
    const float bx1 = box_b[0], by1 = box_b[1], bx2 = box_b[2], by2 = box_b[3], bang = box_b[4];
    P2 ca, cb;
    ca.x = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f); ca.y = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
    cb.x = __fmul_rn(__fadd_rn(bx1, bx2), 0.5f); cb.y = __fmul_rn(__fadd_rn(by1, by2), 0.5f);
    P2 A[5], B[5];
    A[0] = {ax1, ay1}; A[1] = {ax2, ay1}; A[2] = {ax2, ay2}; A[3] = {ax1, ay2};
    B[0] = {bx1, by1}; B[1] = {bx2, by1}; B[2] = {bx2, by2}; B[3] = {bx1, by2};
    const float a_cos = cosf(aang), a_sin = sinf(aang);
    const float b_cos = cosf(bang), b_sin = sinf(bang);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        spin(ca, a_cos, a_sin, A[k]);
        spin(cb, b_cos, b_sin, B[k]);
    }
    A[4] = A[0];
    B[4] = B[0];

    P2 poly[16];
    P2 centre = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersect(A[i + 1], A[i], B[j + 1], B[j], poly[cnt])) {
                centre.x = __fadd_rn(centre.x, poly[cnt].x);
                centre.y = __fadd_rn(centre.y, poly[cnt].y);
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (inside_box(box_a, B[k])) {
            centre.x = __fadd_rn(centre.x, B[k].x);
            centre.y = __fadd_rn(centre.y, B[k].y);
            poly[cnt++] = B[k];
        }
        if (inside_box(box_b, A[k])) {
            centre.x = __fadd_rn(centre.x, A[k].x);
            centre.y = __fadd_rn(centre.y, A[k].y);
            poly[cnt++] = A[k];
        }
    }
    centre.x = __fdiv_rn(centre.x, (float)cnt);
    centre.y = __fdiv_rn(centre.y, (float)cnt);
    // order the vertices by polar angle about the centroid: the same adjacent-swap passes as
    // the reference (:187-196), with each vertex's atan2f evaluated once instead of per compare
    float ang[16];
    for (int i = 0; i < cnt; ++i) ang[i] = atan2f(__fsub_rn(poly[i].y, centre.y), __fsub_rn(poly[i].x, centre.x));
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                const P2 tp = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = tp;
                const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = __fsub_rn(poly[k].x, poly[0].x), uy = __fsub_rn(poly[k].y, poly[0].y);
        const float vx = __fsub_rn(poly[k + 1].x, poly[0].x), vy = __fsub_rn(poly[k + 1].y, poly[0].y);
        area = __fadd_rn(area, diffprod(ux, vy, uy, vx));
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

__device__ __forceinline__ float rot_iou(const float *a, const float *b) {
    const float sb = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    const float s = rot_overlap(a, b);
    const float uni = __fsub_rn(__fmaf_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]), sb), s);
    return __fdiv_rn(s, fmaxf(uni, kEps));
}

// iou3d_kernel.cu:295-303 as compiled inside nms_normal_kernel (the row box's area is the plain product)
__device__ __forceinline__ float flat_iou(const float *a, const float *b) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float uni = __fsub_rn(__fmaf_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]), sa), inter);
    return __fdiv_rn(inter, fmaxf(uni, kEps));
}

template <bool IOU>
__global__ void __launch_bounds__(256) pairwise_kernel(int na, const float *__restrict__ a, int nb,
                                                      const float *__restrict__ b, float *__restrict__ out) {
    // 16 x 16 pairs per CTA, b index fastest so the result rows are written coalesced
    const int ia = blockIdx.y * 16 + threadIdx.y;
    const int ib = blockIdx.x * 16 + threadIdx.x;
    __shared__ float sa[16][5], sb[16][5];
    const int t = threadIdx.y * 16 + threadIdx.x;
    if (t < 80) {
        const int r = t / 5, c = t % 5;
        const int ga = blockIdx.y * 16 + r;
        sa[r][c] = ga < na ? a[ga * 5 + c] : 0.f;
    } else if (t < 160) {
        const int r = (t - 80) / 5, c = (t - 80) % 5;
        const int gb = blockIdx.x * 16 + r;
        sb[r][c] = gb < nb ? b[gb * 5 + c] : 0.f;
    }
    __syncthreads();
    if (ia >= na || ib >= nb) return;
    out[(size_t)ia * nb + ib] = IOU ? rot_iou(sa[threadIdx.y], sb[threadIdx.x]) : rot_overlap(sa[threadIdx.y], sb[threadIdx.x]);
}

// boxes_iou3d_gpu (lib/utils/iou3d/iou3d_utils.py:21-53) in ONE launch: BEV boxes (kitti_utils.py:134-147), rotated
// intersection area, height overlap, volumes and the ratio -- the reference (and the mirror's torch path) spends one kernel
// + ~14 elementwise launches on it, twice per scene in the recall bookkeeping of eval_rcnn.py:541-609.  Float operations
// are torch's, in torch's order: half extents are l / 2 = l * 0.5, clamp(min(a_max, b_max) - max(a_min, b_min), 0),
// (h * w) * l, overlap / clamp(vol_a + vol_b - overlap, 1e-7).
__global__ void __launch_bounds__(256) iou3d_pairs_kernel(int na, const float *__restrict__ a, int nb,
                                                         const float *__restrict__ b, float *__restrict__ out) {
    const int ia = blockIdx.y * 16 + threadIdx.y;
    const int ib = blockIdx.x * 16 + threadIdx.x;
    __shared__ float sa[16][7], sb[16][7];
    const int t = threadIdx.y * 16 + threadIdx.x;
    if (t < 112) {
        const int r = t / 7, c = t % 7;
        const int ga = blockIdx.y * 16 + r;
        sa[r][c] = ga < na ? a[ga * 7 + c] : 0.f;
    } else if (t < 224) {
        const int r = (t - 112) / 7, c = (t - 112) % 7;
        const int gb = blockIdx.x * 16 + r;
        sb[r][c] = gb < nb ? b[gb * 7 + c] : 0.f;
    }
    __syncthreads();
    if (ia >= na || ib >= nb) return;
    const float *pa = sa[threadIdx.y], *pb = sb[threadIdx.x];      // [x, y, z, h, w, l, ry]
    float ba[5], bb[5];
    {
        const float hl = __fmul_rn(pa[5], 0.5f), hw = __fmul_rn(pa[4], 0.5f);
        ba[0] = __fsub_rn(pa[0], hl); ba[1] = __fsub_rn(pa[2], hw); ba[2] = __fadd_rn(pa[0], hl); ba[3] = __fadd_rn(pa[2], hw); ba[4] = pa[6];
    }
    {
        const float hl = __fmul_rn(pb[5], 0.5f), hw = __fmul_rn(pb[4], 0.5f);
        bb[0] = __fsub_rn(pb[0], hl); bb[1] = __fsub_rn(pb[2], hw); bb[2] = __fadd_rn(pb[0], hl); bb[3] = __fadd_rn(pb[2], hw); bb[4] = pb[6];
    }
    const float bev = rot_overlap(ba, bb);
    const float a_min = __fsub_rn(pa[1], pa[3]), b_min = __fsub_rn(pb[1], pb[3]);
    const float oh = fmaxf(__fsub_rn(fminf(pa[1], pb[1]), fmaxf(a_min, b_min)), 0.f);
    const float o3 = __fmul_rn(bev, oh);
    const float va = __fmul_rn(__fmul_rn(pa[3], pa[4]), pa[5]), vb = __fmul_rn(__fmul_rn(pb[3], pb[4]), pb[5]);
    out[(size_t)ia * nb + ib] = __fdiv_rn(o3, fmaxf(__fsub_rn(__fadd_rn(va, vb), o3), 1e-7f));
}

constexpr int kNmsMaxWords = 1024;       // up to 32768 boxes per problem
constexpr int kNmsDense = 128;           // rotated problems up to this size evaluate all pairs up front
constexpr int kNmsStageBytes = 192 * 1024;   // shared-memory budget for the staged boxes of a problem

// One CTA per problem.  boxes (P, stride, 5) sorted by descending score; counts[P] (device) or n.
//  * The boxes of the problem are staged once in shared memory when they fit (9830 boxes): a
//    suppression row then costs shared-memory reads instead of 5 dependent L2 loads per box and
//    step (the lazy rows are latency-bound: 70 kept boxes x n/THREADS steps each).
//  * Small rotated problems (the final NMS of eval_rcnn.py: <= 100 boxes per scene) evaluate all
//    i < j pairs in parallel into a bit matrix first; the greedy pass then only reads bits.  The
//    rotated IoU is ~50x the cost of the axis-aligned one, and the lazy row of a 100-box problem
//    keeps 3 of 8 warps busy for one IoU at a time.
template <bool ROTATED, int THREADS>
__device__ __forceinline__ void nms_body(const int prob, const float *__restrict__ boxes, int stride, int n_fixed,
                                         const int32_t *__restrict__ counts, float thresh, int max_keep,
                                         long long *__restrict__ keep, int32_t *__restrict__ num_out, int stage_cap) {
    extern __shared__ float staged[];     // stage_cap boxes x 5
    __shared__ uint32_t remv[kNmsMaxWords];
    __shared__ int cur;
    __shared__ float cur_box[5];
    const int n = counts ? counts[prob] : n_fixed;
    boxes += (size_t)prob * stride * 5;
    keep += (size_t)prob * max_keep;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int words = (n + 31) >> 5;
    for (int w = tid; w < words; w += THREADS) remv[w] = 0u;
    const bool in_smem = n <= stage_cap;
    if (in_smem)
        for (int t = tid; t < n * 5; t += THREADS) staged[t] = __ldg(boxes + t);
    __syncthreads();
    const float *src = in_smem ? staged : boxes;

    if (ROTATED && n <= kNmsDense && in_smem) {
        // ---- dense mode: sup[i][j] for all i < j, then the greedy pass over bits ----
        uint32_t *sup = remv;             // n rows x wpr words (n <= 128 -> at most 512 words)
        const int wpr = (n + 31) >> 5;
        for (int t = tid; t < n * wpr; t += THREADS) sup[t] = 0u;
        __syncthreads();
        const int pairs = n * (n - 1) / 2;
        for (int t = tid; t < pairs; t += THREADS) {
            // t -> (i, j), i < j, rows enumerated from the last: row i has n-1-i pairs
            int i = (int)((2.0f * n - 1.0f - sqrtf((2.0f * n - 1.0f) * (2.0f * n - 1.0f) - 8.0f * t)) * 0.5f);
            while (i > 0 && i * (2 * n - i - 1) / 2 > t) --i;
            while ((i + 1) * (2 * n - i - 2) / 2 <= t) ++i;
            const int j = t - i * (2 * n - i - 1) / 2 + i + 1;
            if (rot_iou(src + i * 5, src + j * 5) > thresh) atomicOr(&sup[i * wpr + (j >> 5)], 1u << (j & 31));
        }
        __syncthreads();
        if (warp == 0) {
            // lane w owns word w of the running "removed" set (wpr <= 4)
            uint32_t removed = 0u;
            int kept = 0;
            for (int i = 0; i < n && kept < max_keep; ++i) {
                const uint32_t wv = __shfl_sync(0xffffffffu, removed, i >> 5);
                if ((wv >> (i & 31)) & 1u) continue;
                if (lane == 0) keep[kept] = i;
                ++kept;
                if (lane < wpr) removed |= sup[i * wpr + lane];
            }
            if (lane == 0) num_out[prob] = kept;
        }
        return;
    }

    int kept = 0;
    int pos = 0;  // first candidate index not yet examined (uniform)
    while (kept < max_keep && pos < n) {
        // warp 0 finds the first unsuppressed box at index >= pos
        if (warp == 0) {
            int found = n;
            for (int w0 = pos >> 5; w0 < words && found == n; w0 += 32) {
                const int w = w0 + lane;
                uint32_t freeb = 0u;
                if (w < words) {
                    freeb = ~remv[w];
                    if (w == (pos >> 5)) freeb &= 0xffffffffu << (pos & 31);
                    const int hi = n - w * 32;  // valid bits in this word
                    if (hi < 32) freeb &= (1u << hi) - 1u;
                }
                const unsigned any = __ballot_sync(0xffffffffu, freeb != 0u);
                if (any) {
                    const int sl = __ffs(any) - 1;
                    const uint32_t fb = __shfl_sync(0xffffffffu, freeb, sl);
                    found = (w0 + sl) * 32 + (__ffs(fb) - 1);
                }
            }
            if (lane == 0) cur = found;
            if (found < n && lane < 5) cur_box[lane] = src[(size_t)found * 5 + lane];
        }
        __syncthreads();
        const int i = cur;
        if (i >= n) break;
        if (tid == 0) keep[kept] = i;
        ++kept;
        pos = i + 1;
        if (kept < max_keep) {
            // suppression row of box i, only for j > i; one 32-box word per warp step
            float bi[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) bi[c] = cur_box[c];
            for (int w = (pos >> 5) + warp; w < words; w += THREADS / 32) {
                const int j = w * 32 + lane;
                bool sup = false;
                if (j >= pos && j < n && !((remv[w] >> lane) & 1u)) {
                    float bj[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) bj[c] = src[(size_t)j * 5 + c];
                    sup = (ROTATED ? rot_iou(bi, bj) : flat_iou(bi, bj)) > thresh;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sup);
                if (lane == 0 && bal) remv[w] |= bal;
            }
        }
        __syncthreads();
    }
    if (tid == 0) num_out[prob] = kept;
}
static_assert(kNmsDense * (kNmsDense / 32) <= kNmsMaxWords, "dense bit matrix must fit in remv[]");

template <bool ROTATED, int THREADS>
__global__ void __launch_bounds__(THREADS) nms_kernel(const float *__restrict__ boxes, int stride, int n_fixed,
                                                     const int32_t *__restrict__ counts, float thresh, int max_keep,
                                                     long long *__restrict__ keep, int32_t *__restrict__ num_out,
                                                     int stage_cap) {
    nms_body<ROTATED, THREADS>(blockIdx.x, boxes, stride, n_fixed, counts, thresh, max_keep, keep, num_out, stage_cap);
}

// Two independent sets of problems in one launch (the near and the far band of the proposal layer, lib/rpn/proposal_layer.py:
// 58-119: different candidate counts and keep limits): CTAs [0, pa) take set a, the rest set b.  As two launches the second
// band waited for the first although each keeps only 16 SMs busy.
struct NmsSet {
    const float *boxes; int stride; int n; const int32_t *counts; int max_keep; long long *keep; int32_t *num; int cap;
};
template <bool ROTATED, int THREADS>
__global__ void __launch_bounds__(THREADS) nms_pair_kernel(const NmsSet a, const NmsSet b, int pa, float thresh) {
    const bool first = (int)blockIdx.x < pa;
    const NmsSet &s = first ? a : b;
    nms_body<ROTATED, THREADS>(first ? (int)blockIdx.x : (int)blockIdx.x - pa, s.boxes, s.stride, s.n, s.counts, thresh, s.max_keep,
                               s.keep, s.num, s.cap);
}

template <bool ROTATED, int THREADS>
cudaError_t launch_nms(const float *boxes, int problems, int stride, int n, const int32_t *counts, float thresh, int max_keep,
                       long long *keep, int32_t *num, cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(nms_kernel<ROTATED, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsStageBytes);
        attr_done = true;
    }
    int cap = kNmsStageBytes / 20;
    if (cap > stride) cap = stride;
    nms_kernel<ROTATED, THREADS><<<problems, THREADS, (size_t)cap * 20, stream>>>(boxes, stride, n, counts, thresh, max_keep,
                                                                               keep, num, cap);
    return cudaGetLastError();
}

template <bool ROTATED, int THREADS>
cudaError_t launch_nms_pair(NmsSet a, int pa, NmsSet b, int pb, float thresh, cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(nms_pair_kernel<ROTATED, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsStageBytes);
        attr_done = true;
    }
    a.cap = min(kNmsStageBytes / 20, a.stride);
    b.cap = min(kNmsStageBytes / 20, b.stride);
    nms_pair_kernel<ROTATED, THREADS><<<pa + pb, THREADS, (size_t)max(a.cap, b.cap) * 20, stream>>>(a, b, pa, thresh);
    return cudaGetLastError();
}

}  // namespace

// boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap)  iou3d.cpp:31-50 / iou3d_kernel.cu:223-234.
// a (na,5), b (nb,5) [x1,y1,x2,y2,ry] -> out (na,nb) intersection AREA.
PN2_API int pn2_boxes_overlap_bev_f32(const float *a, int na, const float *b, int nb, float *out, cudaStream_t stream) {
    if (na < 0 || nb < 0) { pn2_set_last_error("pn2_boxes_overlap_bev_f32: bad argument"); return PN2_ERR_INVALID; }
    if (na == 0 || nb == 0) return PN2_OK;
    dim3 grid(pn2_divup(nb, 16), pn2_divup(na, 16)), block(16, 16);
    pairwise_kernel<false><<<grid, block, 0, stream>>>(na, a, nb, b, out);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// boxes_iou_bev_gpu  iou3d.cpp:52-71 / iou3d_kernel.cu:236-248.
PN2_API int pn2_boxes_iou_bev_f32(const float *a, int na, const float *b, int nb, float *out, cudaStream_t stream) {
    if (na < 0 || nb < 0) { pn2_set_last_error("pn2_boxes_iou_bev_f32: bad argument"); return PN2_ERR_INVALID; }
    if (na == 0 || nb == 0) return PN2_OK;
    dim3 grid(pn2_divup(nb, 16), pn2_divup(na, 16)), block(16, 16);
    pairwise_kernel<true><<<grid, block, 0, stream>>>(na, a, nb, b, out);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// boxes_iou3d_gpu (lib/utils/iou3d/iou3d_utils.py:21-53): a (na,7), b (nb,7) [x,y,z,h,w,l,ry] -> out (na,nb) 3-D IoU.
PN2_API int pn2_boxes_iou3d_f32(const float *a, int na, const float *b, int nb, float *out, cudaStream_t stream) {
    if (na < 0 || nb < 0 || (na > 0 && nb > 0 && (!a || !b || !out))) { pn2_set_last_error("pn2_boxes_iou3d_f32: bad argument"); return PN2_ERR_INVALID; }
    if (na == 0 || nb == 0) return PN2_OK;
    dim3 grid(pn2_divup(nb, 16), pn2_divup(na, 16)), block(16, 16);
    iou3d_pairs_kernel<<<grid, block, 0, stream>>>(na, a, nb, b, out);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// nms_gpu / nms_normal_gpu (iou3d.cpp:73-169, kernels iou3d_kernel.cu:250-348), batched and
// entirely on the device.  boxes (problems, stride, 5) sorted by descending score; problem p
// uses its first counts[p] boxes (counts may be NULL: all use n).  keep (problems, max_keep)
// int64 indices into the sorted list in ascending order; num (problems) int32 survivors
// (<= max_keep; pass max_keep = n for the reference's full list).
PN2_API int pn2_nms_bev_f32(const float *boxes, int problems, int stride, int n, const int32_t *counts, float thresh,
                            int rotated, int max_keep, long long *keep, int32_t *num, cudaStream_t stream) {
    if (problems < 0 || stride < 0 || n < 0 || n > stride || max_keep < 0 || stride > kNmsMaxWords * 32) {
        pn2_set_last_error("pn2_nms_bev_f32: bad argument (at most 32768 boxes per problem)");
        return PN2_ERR_INVALID;
    }
    if (problems == 0) return PN2_OK;
    const cudaError_t e = rotated ? launch_nms<true, 256>(boxes, problems, stride, n, counts, thresh, max_keep, keep, num, stream)
                                  : launch_nms<false, 1024>(boxes, problems, stride, n, counts, thresh, max_keep, keep, num, stream);
    if (e != cudaSuccess) {
        pn2_set_last_error(cudaGetErrorString(e));
        return PN2_ERR_LAUNCH;
    }
    return PN2_OK;
}

// pn2_nms_bev_f32 for two sets of problems in ONE launch (same results as two calls): set 0 = problems0 x (stride0, n0,
// counts0, max_keep0) -> keep0 / num0, set 1 likewise; one threshold and one NMS type for both.
PN2_API int pn2_nms_bev_pair_f32(const float *boxes0, int problems0, int stride0, int n0, const int32_t *counts0, int max_keep0,
                                 long long *keep0, int32_t *num0, const float *boxes1, int problems1, int stride1, int n1,
                                 const int32_t *counts1, int max_keep1, long long *keep1, int32_t *num1, float thresh,
                                 int rotated, cudaStream_t stream) {
    if (problems0 < 0 || problems1 < 0 || stride0 < 0 || stride1 < 0 || n0 < 0 || n1 < 0 || n0 > stride0 || n1 > stride1 ||
        max_keep0 < 0 || max_keep1 < 0 || stride0 > kNmsMaxWords * 32 || stride1 > kNmsMaxWords * 32) {
        pn2_set_last_error("pn2_nms_bev_pair_f32: bad argument (at most 32768 boxes per problem)");
        return PN2_ERR_INVALID;
    }
    if (problems0 + problems1 == 0) return PN2_OK;
    const NmsSet a = {boxes0, stride0, n0, counts0, max_keep0, keep0, num0, 0};
    const NmsSet b = {boxes1, stride1, n1, counts1, max_keep1, keep1, num1, 0};
    const cudaError_t e = rotated ? launch_nms_pair<true, 256>(a, problems0, b, problems1, thresh, stream)
                                  : launch_nms_pair<false, 1024>(a, problems0, b, problems1, thresh, stream);
    if (e != cudaSuccess) {
        pn2_set_last_error(cudaGetErrorString(e));
        return PN2_ERR_LAUNCH;
    }
    return PN2_OK;
}
