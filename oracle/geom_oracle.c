/*
 * geom_oracle.c -- CPU restatement of the reference's rotated-box geometry:
 *   pointrcnn/lib/utils/iou3d/src/iou3d_kernel.cu  (box_overlap, iou_bev, nms_kernel)
 *   pointrcnn/lib/utils/iou3d/src/iou3d.cpp:73-119 (host greedy pass over the masks)
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
 * --impl reference).  The product never links this file.
 *
 * Float expressions use explicit fmaf() exactly where the COMPILED reference fuses (nvcc front
 * end + ptxas; read from the SASS of oracle/_ref/libpn2_legacy.so): a*b -+ c*d runs as
 * fma(a, b, -+rn(c*d)) everywhere except the s2 / s5 pair of a segment test, whose two shared
 * products are rounded first.  Build with -ffp-contract=off.  sinf/cosf/atan2f come from the host libm, which can differ from the
 * CUDA math library in the last ulp: the pin (tests/test_golden_cpu.py) therefore checks
 * areas against the reference-kernel goldens to 5e-5 m^2 (the shoelace sum amplifies the trig ulp),
 * the zero pattern and the NMS keep lists exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } Pt;

/* a*b - c*d as the compiled reference executes it (SASS of the unmodified kernel): the minuend
 * product is fused, the other product is rounded first */
static inline float diffprod(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }

static const float EPS = 1e-8f;

/* iou3d_kernel.cu:38-40 cross(p1, p2, p0) */
static inline float cross3(Pt p1, Pt p2, Pt p0) {
    return diffprod(p1.x - p0.x, p2.y - p0.y, p2.x - p0.x, p1.y - p0.y);
}

/* iou3d_kernel.cu:42-48 */
static inline int check_rect_cross(Pt p1, Pt p2, Pt q1, Pt q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

/* iou3d_kernel.cu:50-64 */
static inline int check_in_box2d(const float *box, Pt p) {
    const float MARGIN = 1e-5f;
    const float center_x = (box[0] + box[2]) / 2, center_y = (box[1] + box[3]) / 2;
    const float angle_cos = cosf(-box[4]), angle_sin = sinf(-box[4]);
    const float dx = p.x - center_x, dy = p.y - center_y;
    const float rot_x = fmaf(dx, angle_cos, dy * angle_sin) + center_x;
    const float rot_y = diffprod(angle_cos, dy, angle_sin, dx) + center_y;
    return rot_x > box[0] - MARGIN && rot_x < box[2] + MARGIN && rot_y > box[1] - MARGIN && rot_y < box[3] + MARGIN;
}

/* iou3d_kernel.cu:66-96 */
static inline int intersection(Pt p1, Pt p0, Pt q1, Pt q0, Pt *ans) {
    if (!check_rect_cross(p0, p1, q0, q1)) return 0;
    /* s2 and s5 = -s2 share their two products: both rounded, then subtracted */
    const float pa = (p1.x - p0.x) * (q1.y - p0.y), pb = (q1.x - p0.x) * (p1.y - p0.y);
    const float s1 = cross3(q0, p1, p0), s2 = pa - pb, s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = pb - pa;
    if (fabsf(s5 - s1) > EPS) {
        ans->x = diffprod(s5, q0.x, s1, q1.x) / (s5 - s1);
        ans->y = diffprod(s5, q0.y, s1, q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = diffprod(p0.x, p1.y, p1.x, p0.y);
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = diffprod(q0.x, q1.y, q1.x, q0.y);
        const float D = diffprod(a0, b1, a1, b0);
        ans->x = diffprod(b0, c1, b1, c0) / D;
        ans->y = diffprod(a1, c0, a0, c1) / D;
    }
    return 1;
}

/* iou3d_kernel.cu:98-102 */
static inline Pt rotate_around_center(Pt c, float angle_cos, float angle_sin, Pt p) {
    const float dx = p.x - c.x, dy = p.y - c.y;
    Pt r;
    r.x = fmaf(dx, angle_cos, dy * angle_sin) + c.x;
    r.y = diffprod(angle_cos, dy, angle_sin, dx) + c.y;
    return r;
}

/* iou3d_kernel.cu:108-212 box_overlap */
float orc_box_overlap(const float *box_a, const float *box_b) {
    const float a_x1 = box_a[0], a_y1 = box_a[1], a_x2 = box_a[2], a_y2 = box_a[3], a_angle = box_a[4];
    const float b_x1 = box_b[0], b_y1 = box_b[1], b_x2 = box_b[2], b_y2 = box_b[3], b_angle = box_b[4];
    Pt ca = {(a_x1 + a_x2) / 2, (a_y1 + a_y2) / 2}, cb = {(b_x1 + b_x2) / 2, (b_y1 + b_y2) / 2};
    Pt A[5] = {{a_x1, a_y1}, {a_x2, a_y1}, {a_x2, a_y2}, {a_x1, a_y2}};
    Pt B[5] = {{b_x1, b_y1}, {b_x2, b_y1}, {b_x2, b_y2}, {b_x1, b_y2}};
    const float a_cos = cosf(a_angle), a_sin = sinf(a_angle), b_cos = cosf(b_angle), b_sin = sinf(b_angle);
    for (int k = 0; k < 4; ++k) {
        A[k] = rotate_around_center(ca, a_cos, a_sin, A[k]);
        B[k] = rotate_around_center(cb, b_cos, b_sin, B[k]);
    }
    A[4] = A[0];
    B[4] = B[0];
    Pt cp[16], centre = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (intersection(A[i + 1], A[i], B[j + 1], B[j], &cp[cnt])) {
                centre.x = centre.x + cp[cnt].x;
                centre.y = centre.y + cp[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (check_in_box2d(box_a, B[k])) {
            centre.x = centre.x + B[k].x; centre.y = centre.y + B[k].y;
            cp[cnt++] = B[k];
        }
        if (check_in_box2d(box_b, A[k])) {
            centre.x = centre.x + A[k].x; centre.y = centre.y + A[k].y;
            cp[cnt++] = A[k];
        }
    }
    centre.x /= cnt;
    centre.y /= cnt;
    /* :187-196 adjacent-swap passes on the polar angle about the centroid */
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(cp[i].y - centre.y, cp[i].x - centre.x) > atan2f(cp[i + 1].y - centre.y, cp[i + 1].x - centre.x)) {
                const Pt t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = cp[k].x - cp[0].x, uy = cp[k].y - cp[0].y;
        const float vx = cp[k + 1].x - cp[0].x, vy = cp[k + 1].y - cp[0].y;
        area += diffprod(ux, vy, uy, vx);
    }
    return (float)(fabs((double)area) / 2.0);
}

/* iou3d_kernel.cu:214-221 iou_bev */
float orc_iou_bev(const float *a, const float *b) {
    const float sb = (b[2] - b[0]) * (b[3] - b[1]);
    const float s = orc_box_overlap(a, b);
    return s / fmaxf(fmaf(a[2] - a[0], a[3] - a[1], sb) - s, EPS);
}

/* boxes_overlap_kernel :223-234 / boxes_iou_bev_kernel :236-248 */
void orc_boxes_overlap_bev(const float *a, int na, const float *b, int nb, float *out, int iou) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = iou ? orc_iou_bev(a + i * 5, b + j * 5) : orc_box_overlap(a + i * 5, b + j * 5);
}

/* nms_kernel :250-292 + iou3d.cpp:100-116.  boxes sorted by descending score.  Greedy pass:
 * box i is kept iff no kept j < i has iou(j, i) > thresh (row box = the kept one, like the
 * mask kernel where the row index is the suppressor). */
int orc_nms_rotated(const float *boxes, int64_t *keep, int n, float thresh) {
    uint8_t *dead = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[kept++] = i;
        for (int j = i + 1; j < n; ++j)
            if (!dead[j] && orc_iou_bev(boxes + (size_t)i * 5, boxes + (size_t)j * 5) > thresh) dead[j] = 1;
    }
    free(dead);
    return kept;
}

/* iou3d.cpp:100-116 (and :150-166): the host greedy pass over the u64 suppression masks the
 * reference copies back from the device.  mask (n, col_blocks) u64 -> keep, returns the count. */
int orc_nms_greedy_from_mask(const unsigned long long *mask, int n, int64_t *keep) {
    const int col_blocks = n / 64 + (n % 64 > 0);
    unsigned long long *remv = (unsigned long long *)calloc((size_t)(col_blocks > 0 ? col_blocks : 1), sizeof(unsigned long long));
    int num_to_keep = 0;
    for (int i = 0; i < n; ++i) {
        const int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ULL << inblock))) {
            keep[num_to_keep++] = i;
            const unsigned long long *p = mask + (size_t)i * col_blocks;
            for (int j = nblock; j < col_blocks; ++j) remv[j] |= p[j];
        }
    }
    free(remv);
    return num_to_keep;
}

/* ------------------------------------------------------------------------------------------
 * evaluate/rotate_iou.py:16-291 (numba.cuda kernel rotate_iou_kernel_eval and device functions),
 * restated with the float32 / float64 mix numba infers and the FMA fusions of the compiled
 * kernel (PTX from numba 0.65 on a B200 + the SASS ptxas 12.9 makes of it; see
 * csrc/rotate_iou.cu for the list).  npts_out (optional) receives the number of polygon
 * vertices per pair: the reference keeps them in a local array of 8 points (rotate_iou.py:233),
 * so pairs with more than 8 are out-of-bounds writes there -- undefined, not reproducible.
 * ------------------------------------------------------------------------------------------ */
static void riou_corners(const float *rb, float *c) {            /* :203-227 */
    const float a_cos = cosf(rb[4]), a_sin = sinf(rb[4]);
    const float xh = rb[2] * 0.5f, yh = rb[3] * 0.5f;
    const float px[4] = {-xh, -xh, xh, xh}, py[4] = {-yh, yh, yh, -yh};
    for (int i = 0; i < 4; ++i) {
        c[2 * i] = fmaf(px[i], a_cos, a_sin * py[i]) + rb[0];
        c[2 * i + 1] = fmaf(py[i], a_cos, -(px[i] * a_sin)) + rb[1];
    }
}
static int riou_point_in_quad(float x, float y, const float *q) { /* :160-176 */
    const float ab0 = q[2] - q[0], ab1 = q[3] - q[1], ad0 = q[6] - q[0], ad1 = q[7] - q[1];
    const float ap0 = x - q[0], ap1 = y - q[1];
    const float abab = fmaf(ab0, ab0, ab1 * ab1), abap = fmaf(ab1, ap1, ab0 * ap0);
    const float adad = fmaf(ad0, ad0, ad1 * ad1), adap = fmaf(ad1, ap1, ad0 * ap0);
    return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}
static int riou_segment(const float *p1, const float *p2, int i, int j, float *out) { /* :72-115 */
    const float A0 = p1[2 * i], A1 = p1[2 * i + 1], B0 = p1[2 * ((i + 1) % 4)], B1 = p1[2 * ((i + 1) % 4) + 1];
    const float C0 = p2[2 * j], C1 = p2[2 * j + 1], D0 = p2[2 * ((j + 1) % 4)], D1 = p2[2 * ((j + 1) % 4) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const int acd = DA1 * CA0 > CA1 * DA0;
    const int bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
    if (acd == bcd) return 0;
    const int abc = CA1 * BA0 > BA1 * CA0, abd = DA1 * BA0 > BA1 * DA0;
    if (abc == abd) return 0;
    const float DC0 = D0 - C0, DC1 = D1 - C1;
    const float ABBA = fmaf(A0, B1, -(B0 * A1)), CDDC = fmaf(C0, D1, -(D0 * C1));
    const float DH = fmaf(BA1, DC0, -(BA0 * DC1));
    out[0] = fmaf(ABBA, DC0, -(BA0 * CDDC)) / DH;
    out[1] = fmaf(ABBA, DC1, -(BA1 * CDDC)) / DH;
    return 1;
}
static double riou_inter(const float *r1, const float *r2, int *npts) {  /* :230-244 */
    float c1[8], c2[8], pts[48], vs[24];
    riou_corners(r1, c1);
    riou_corners(r2, c2);
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        if (riou_point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) { pts[2 * n] = c1[2 * i]; pts[2 * n + 1] = c1[2 * i + 1]; ++n; }
        if (riou_point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) { pts[2 * n] = c2[2 * i]; pts[2 * n + 1] = c2[2 * i + 1]; ++n; }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float t[2];
            if (riou_segment(c1, c2, i, j, t)) { pts[2 * n] = t[0]; pts[2 * n + 1] = t[1]; ++n; }
        }
    if (npts) *npts = n;
    if (n > 0) {                                                      /* :32-69 */
        float s0 = 0.f, s1 = 0.f;
        for (int i = 0; i < n; ++i) { s0 += pts[2 * i]; s1 += pts[2 * i + 1]; }
        const float m0 = (float)((double)s0 / (double)n), m1 = (float)((double)s1 / (double)n);
        for (int i = 0; i < n; ++i) {
            float v0 = pts[2 * i] - m0, v1 = pts[2 * i + 1] - m1;
            const float d = sqrtf(fmaf(v0, v0, v1 * v1));
            v0 = v0 / d; v1 = v1 / d;
            if (v1 < 0.f) v0 = -2.f - v0;
            vs[i] = v0;
        }
        for (int i = 1; i < n; ++i)
            if (vs[i - 1] > vs[i]) {
                const float temp = vs[i], tx = pts[2 * i], ty = pts[2 * i + 1];
                int j = i;
                while (j > 0 && vs[j - 1] > temp) { vs[j] = vs[j - 1]; pts[2 * j] = pts[2 * j - 2]; pts[2 * j + 1] = pts[2 * j - 1]; --j; }
                vs[j] = temp; pts[2 * j] = tx; pts[2 * j + 1] = ty;
            }
    }
    double area = 0.0;                                                /* :16-29 */
    for (int i = 0; i < n - 2; ++i) {
        const float *a = pts, *b = pts + 2 * i + 2, *c = pts + 2 * i + 4;
        const float cr = fmaf(a[0] - c[0], b[1] - c[1], -((a[1] - c[1]) * (b[0] - c[0])));
        area += fabs((double)cr * 0.5);
    }
    return area;
}
/* :247-291.  boxes (n,5), qboxes (k,5) -> out (n,k); rbox1 = query box, rbox2 = box */
void orc_rotate_iou_eval(const float *boxes, int n, const float *qboxes, int k, float *out, int criterion, int32_t *npts_out) {
    for (int ib = 0; ib < n; ++ib)
        for (int iq = 0; iq < k; ++iq) {
            const float *r1 = qboxes + (size_t)iq * 5, *r2 = boxes + (size_t)ib * 5;
            const float area1 = r1[2] * r1[3], area2 = r2[2] * r2[3];
            int np = 0;
            const double ai = riou_inter(r1, r2, &np);
            double r;
            if (criterion == -1) r = ai / ((double)(area1 + area2) - ai);
            else if (criterion == 0) r = ai / (double)area1;
            else if (criterion == 1) r = ai / (double)area2;
            else r = ai;
            out[(size_t)ib * k + iq] = (float)r;
            if (npts_out) npts_out[(size_t)ib * k + iq] = np;
        }
}
