/*
 * pn2_oracle.c -- CPU restatement of the reference's pointnet2 / roipool3d / iou3d-normal
 * kernels.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product path never links this.
 *
 * Every function restates one reference CUDA kernel as a scalar loop nest and cites the
 * file:line it follows (paths relative to /root/reference/).  Floating-point expressions are
 * written with explicit fmaf() in the contraction order nvcc 12.9 emits for the reference
 * source (read from the PTX of the unmodified .cu files; see DESIGN.md "FMA contraction").
 * Build with -ffp-contract=off so the C compiler adds no contractions of its own.
 *
 * Parity pin: validated on the GPU box against oracle/_ref/libpn2_legacy.so, i.e. the
 * reference's own .cu files compiled unchanged (tests/test_legacy_parity.py), and against
 * the golden vectors under tests/golden/ that were produced by those kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pointrcnn/pointnet2_lib/pointnet2/src/cuda_utils.h:10-14 : block size the reference FPS
 * launcher picks; it decides the tie-break order of the tree reduction. */
int orc_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/* squared distance in the reference's contraction order:
 * t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t)
 * (sampling_gpu.cu:134, ball_query_gpu.cu:33, interpolate_gpu.cu:37 as compiled) */
static inline float sqdist(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    t = fmaf(dz, dz, t);
    return t;
}

/* sampling_gpu.cu:93-209 furthest_point_sampling_kernel<block_size>, one cloud per block.
 * xyz (B,N,3) ; temp (B,N) caller scratch (pre-filled, mutated) ; idx (B,M) int32.
 * The per-thread strided scan (strict >, start best=-1/besti=0) and the shared-memory
 * tree (__update keeps the lower slot on ties, :86-91) are simulated literally. */
void orc_fps(const float *xyz, float *temp, int32_t *idx, int b, int n, int m) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int32_t *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    float d = sqdist(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
                    float d2 = fminf(d, t[k]);
                    t[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = dists[tid], v2 = dists[tid + s];
                    int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2; /* max(v1,v2) */
                    if (v2 != v2 && v1 == v1) dists[tid] = v1;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* sampling_gpu.cu:8-24 gather_points_kernel_fast: out[b,c,m] = points[b,c,idx[b,m]] */
void orc_gather_points(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int mi = 0; mi < m; ++mi)
                out[((size_t)bi * c + ci) * m + mi] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + mi]];
}

/* sampling_gpu.cu:46-63 gather_points_grad_kernel_fast (atomicAdd order is unspecified in
 * the reference; here it is ascending m) */
void orc_gather_points_grad(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c, int n, int m) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int mi = 0; mi < m; ++mi)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + mi]] += grad_out[((size_t)bi * c + ci) * m + mi];
}

/* ball_query_gpu.cu:9-45 ball_query_kernel_fast.  idx rows with no hit are NOT written
 * (the Python caller zero-initialises, pointnet2_utils.py:218). */
void orc_ball_query(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n, int m, float radius, int nsample) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        for (int mi = 0; mi < m; ++mi) {
            const float *c = new_xyz + ((size_t)bi * m + mi) * 3;
            int32_t *row = idx + ((size_t)bi * m + mi) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist(c[0] - p[k * 3 + 0], c[1] - p[k * 3 + 1], c[2] - p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    }
}

/* group_points_gpu.cu:47-66 group_points_kernel_fast: out[b,c,m,s] = points[b,c,idx[b,m,s]] */
void orc_group_points(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m, int ns) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            for (size_t e = 0; e < (size_t)m * ns; ++e)
                out[((size_t)bi * c + ci) * m * ns + e] = src[idx[(size_t)bi * m * ns + e]];
        }
}

/* group_points_gpu.cu:8-25 group_points_grad_kernel_fast */
void orc_group_points_grad(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c, int n, int m, int ns) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            for (size_t e = 0; e < (size_t)m * ns; ++e)
                dst[idx[(size_t)bi * m * ns + e]] += grad_out[((size_t)bi * c + ci) * m * ns + e];
        }
}

/* interpolate_gpu.cu:9-52 three_nn_kernel_fast.  best* are double initialised to 1e40 and
 * compared against the float distance (promoted), strict <, so the lowest index wins ties;
 * dist2 is the double cast back to float (inf if never replaced). */
void orc_three_nn(const float *unknown, const float *known, float *dist2, int32_t *idx, int b, int n, int m) {
    for (int bi = 0; bi < b; ++bi) {
        const float *kn = known + (size_t)bi * m * 3;
        for (int i = 0; i < n; ++i) {
            const float *u = unknown + ((size_t)bi * n + i) * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int b1 = 0, b2 = 0, b3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist(u[0] - kn[k * 3 + 0], u[1] - kn[k * 3 + 1], u[2] - kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
                } else if (d < best2) {
                    best3 = best2; b3 = b2; best2 = d; b2 = k;
                } else if (d < best3) {
                    best3 = d; b3 = k;
                }
            }
            float *dout = dist2 + ((size_t)bi * n + i) * 3;
            int32_t *iout = idx + ((size_t)bi * n + i) * 3;
            dout[0] = (float)best1; dout[1] = (float)best2; dout[2] = (float)best3;
            iout[0] = b1; iout[1] = b2; iout[2] = b3;
        }
    }
}

/* interpolate_gpu.cu:77-97 three_interpolate_kernel_fast; contraction as compiled:
 * t = w1*p1 ; t = fma(w0,p0,t) ; t = fma(w2,p2,t) */
void orc_three_interpolate(const float *points, const int32_t *idx, const float *weight, float *out, int b, int c, int m, int n) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; ++i) {
                const float *w = weight + ((size_t)bi * n + i) * 3;
                const int32_t *id = idx + ((size_t)bi * n + i) * 3;
                float t = w[1] * src[id[1]];
                t = fmaf(w[0], src[id[0]], t);
                t = fmaf(w[2], src[id[2]], t);
                out[((size_t)bi * c + ci) * n + i] = t;
            }
        }
}

/* interpolate_gpu.cu:120-142 three_interpolate_grad_kernel_fast */
void orc_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *weight, float *grad_points, int b, int c, int n, int m) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; ++i) {
                const float *w = weight + ((size_t)bi * n + i) * 3;
                const int32_t *id = idx + ((size_t)bi * n + i) * 3;
                float g = grad_out[((size_t)bi * c + ci) * n + i];
                dst[id[0]] += g * w[0];
                dst[id[1]] += g * w[1];
                dst[id[2]] += g * w[2];
            }
        }
}

/* iou3d_kernel.cu:295-303 iou_normal (axis-aligned BEV IoU, ry ignored), EPS = 1e-8 (:13).
 * As compiled: interS = w*h ; Sa, Sb plain products ; (Sa + Sb) - interS ; div.rn. */
static inline float iou_normal(const float *a, const float *b) {
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    float interS = width * height;
    float Sa = (a[2] - a[0]) * (a[3] - a[1]);
    float Sb = (b[2] - b[0]) * (b[3] - b[1]);
    return interS / fmaxf(Sa + Sb - interS, 1e-8f);
}

/* iou3d_kernel.cu:306-348 nms_normal_kernel (64x64 suppression bit tiles) followed by the
 * host greedy pass of iou3d.cpp:137-169.  boxes (n,5) sorted by score; keep (n) int64;
 * returns num_to_keep. */
int orc_nms_normal(const float *boxes, int64_t *keep, int n, float thresh) {
    const int cb = (n + 63) / 64;
    uint64_t *mask = (uint64_t *)calloc((size_t)n * cb + 1, sizeof(uint64_t));
    uint64_t *remv = (uint64_t *)calloc((size_t)cb + 1, sizeof(uint64_t));
    for (int i = 0; i < n; ++i)
        for (int cbk = 0; cbk < cb; ++cbk) {
            int cs = n - cbk * 64 < 64 ? n - cbk * 64 : 64;
            int start = (i / 64 == cbk) ? (i % 64) + 1 : 0;
            uint64_t t = 0;
            for (int j = start; j < cs; ++j)
                if (iou_normal(boxes + (size_t)i * 5, boxes + (size_t)(cbk * 64 + j) * 5) > thresh) t |= 1ULL << j;
            mask[(size_t)i * cb + cbk] = t;
        }
    int num = 0;
    for (int i = 0; i < n; ++i) {
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {
            keep[num++] = i;
            for (int j = nb; j < cb; ++j) remv[j] |= mask[(size_t)i * cb + j];
        }
    }
    free(mask);
    free(remv);
    return num;
}

/* roipool3d_kernel.cu:14-28 pt_in_box3d.  As compiled: cy and the half extents are formed
 * in double from float operands (h/2.0 etc.), fabsf differences are float promoted for the
 * compare, cos/sin are the precise float versions.  cosa/sina are passed in so the caller
 * can supply either libm or CUDA-libdevice values (they differ in the last ulp). */
static inline int pt_in_box3d(float x, float y, float z, float cx, float by, float cz, float h, float w, float l,
                              float cosa, float sina, float max_dis) {
    float cy = (float)((double)by - (double)h / 2.0);
    if ((fabsf(x - cx) > max_dis) || ((double)fabsf(y - cy) > (double)h / 2.0) || (fabsf(z - cz) > max_dis)) return 0;
    /* as compiled (SASS of the unmodified reference, nvcc/ptxas 12.9): ptxas fuses the PTX
     * mul/sub pair, so  x_rot = fma(dx, cos, -fl(dz*sin)) ;  z_rot = fma(dz, cos, fl(dx*sin)) */
    float x_rot = fmaf(x - cx, cosa, -((z - cz) * sina));
    float z_rot = fmaf(z - cz, cosa, (x - cx) * sina);
    return ((double)x_rot >= -(double)l / 2.0) & ((double)x_rot <= (double)l / 2.0) &
           ((double)z_rot >= -(double)w / 2.0) & ((double)z_rot <= (double)w / 2.0);
}

/* roipool3d_kernel.cu:97-194 assign_pts_to_box3d + get_pooled_idx + roipool3d_forward(idx).
 * xyz (B,N,3) ; boxes3d (B,M,7) already enlarged ; feat (B,N,C) ; pooled (B,M,S,3+C)
 * pre-zeroed ; empty (B,M) pre-zeroed.  trig (B,M,2) optional precomputed cos/sin (NULL =
 * use libm cosf/sinf). */
void orc_roipool3d(const float *xyz, const float *boxes3d, const float *feat, float *pooled, int32_t *empty,
                   const float *trig, int b, int n, int m, int c, int s) {
    int32_t *pidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)s);
    for (int bi = 0; bi < b; ++bi)
        for (int mi = 0; mi < m; ++mi) {
            const float *bx = boxes3d + ((size_t)bi * m + mi) * 7;
            float cosa = trig ? trig[((size_t)bi * m + mi) * 2 + 0] : cosf(bx[6]);
            float sina = trig ? trig[((size_t)bi * m + mi) * 2 + 1] : sinf(bx[6]);
            int cnt = 0;
            for (int k = 0; k < n && cnt < s; ++k) {
                const float *p = xyz + ((size_t)bi * n + k) * 3;
                if (pt_in_box3d(p[0], p[1], p[2], bx[0], bx[1], bx[2], bx[3], bx[4], bx[5], cosa, sina, 10.0f))
                    pidx[cnt++] = k;
            }
            if (cnt == 0) { empty[(size_t)bi * m + mi] = 1; continue; }
            for (int k = cnt; k < s; ++k) pidx[k] = pidx[k % cnt];
            for (int k = 0; k < s; ++k) {
                float *dst = pooled + (((size_t)bi * m + mi) * s + k) * (3 + c);
                const float *p = xyz + ((size_t)bi * n + pidx[k]) * 3;
                dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
                memcpy(dst + 3, feat + ((size_t)bi * n + pidx[k]) * c, sizeof(float) * (size_t)c);
            }
        }
    free(pidx);
}
