// glue.cu -- the small stages between the big kernels of one PointRCNN inference step, each written as ONE kernel
// instead of the 20-60 elementwise / gather / scatter / cat launches the reference's torch statements turn into:
//
//   pn2_decode_bbox_f32          lib/utils/bbox_transform.py:24-121 decode_bbox_target (+ proposal_layer.py:23 for the RPN)
//   pn2_proposal_select_f32      lib/rpn/proposal_layer.py:58-100 distance_based_proposal up to the NMS input (both bands)
//   pn2_proposal_assemble_f32    lib/rpn/proposal_layer.py:107-119 keep lists -> zero-padded (B, 100, 7) ROIs + scores
//   pn2_rcnn_post_prepare_f32    tools/eval_rcnn.py:516-535, 611-620 decode, sigmoid threshold, score order, BEV boxes
//   pn2_rcnn_post_assemble_f32   tools/eval_rcnn.py:621-627 keep list -> fixed-width detection records
//
// Every float operation is the IEEE single-precision operation torch executes for the corresponding statement, in the
// same order, written with explicit _rn intrinsics so that nothing is contracted into an FMA: the decoded boxes feed
// thresholds and NMS and must not move by an ulp.  Python scalars arrive as doubles and are rounded to float exactly
// where torch rounds them (at the elementwise op).  The one operation whose rounding is a property of a library
// kernel -- the K = 2 batched matmul of rotate_pc_along_y_torch (bbox_transform.py:19-20) -- is selectable
// (rot_mode) and pinned by tests/test_glue_gpu.py against torch on the box.
#include "common.cuh"
#include <cfloat>

namespace {

struct DecodeParams {
    const float *roi; int roi_dim;          // (rows, 3) point xyz or (rows, 7) roi
    const float *reg; int c;                // (rows, c)
    float *out;                             // (rows, 7) [x, y, z, h, w, l, ry]
    long long rows;
    int nb; float lbs, lbs_half, loc_scope; int xz_fine;
    int y_by_bin, nby; float ybs, ybs_half, y_scope;
    int nhb, ry_fine; float apc, apc_half, quarter_pi, two_pi, pi;
    float anchor[3];
    int y_bottom;                           // proposals[:, 1] += proposals[:, 3] / 2  (proposal_layer.py:23)
    int rot_mode;
};

// torch.argmax over `n` floats: the first maximum; a NaN counts as the maximum (ATen ArgMaxOps)
__device__ __forceinline__ int argmax_first(const float *v, int n) {
    int best = 0;
    float bv = v[0];
    for (int i = 1; i < n; ++i) {
        const float x = v[i];
        if (!(bv != bv) && (x > bv || x != x)) { bv = x; best = i; }
    }
    return best;
}

// a0 * b0 + a1 * b1 the way the batched K = 2 matmul rounds it
__device__ __forceinline__ float dot2(float a0, float b0, float a1, float b1, int mode) {
    if (mode == 0) return __fmaf_rn(a1, b1, __fmul_rn(a0, b0));
    if (mode == 1) return __fmaf_rn(a0, b0, __fmul_rn(a1, b1));
    return __fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}

// torch.remainder(a, b) for floats (ATen: fmod, then the sign fix)
__device__ __forceinline__ float remainder_torch(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = __fadd_rn(m, b);
    return m;
}

__device__ __forceinline__ void decode_row(const DecodeParams &p, const float *reg, const float *roi, float *o) {
    const int nb = p.nb;
    const int xb = argmax_first(reg, nb), zb = argmax_first(reg + nb, nb);
    float pos_x = __fadd_rn(__fadd_rn(__fmul_rn((float)xb, p.lbs), p.lbs_half), -p.loc_scope);
    float pos_z = __fadd_rn(__fadd_rn(__fmul_rn((float)zb, p.lbs), p.lbs_half), -p.loc_scope);
    int off = nb * 2;
    if (p.xz_fine) {
        pos_x = __fadd_rn(pos_x, __fmul_rn(reg[nb * 2 + xb], p.lbs));
        pos_z = __fadd_rn(pos_z, __fmul_rn(reg[nb * 3 + zb], p.lbs));
        off = nb * 4;
    }
    float pos_y;
    if (p.y_by_bin) {
        const int yb = argmax_first(reg + off, p.nby);
        const float yr = reg[off + p.nby + yb];
        pos_y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)yb, p.ybs), p.ybs_half), -p.y_scope), __fmul_rn(yr, p.ybs));
        pos_y = __fadd_rn(pos_y, roi[1]);
        off += 2 * p.nby;
    } else {
        pos_y = __fadd_rn(roi[1], reg[off]);
        off += 1;
    }
    const int rb = argmax_first(reg + off, p.nhb);
    const float ry_res = __fmul_rn(reg[off + p.nhb + rb], p.apc_half);
    float ry;
    if (p.ry_fine) {
        ry = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)rb, p.apc), p.apc_half), ry_res), -p.quarter_pi);
    } else {
        ry = remainder_torch(__fadd_rn(__fmul_rn((float)rb, p.apc), ry_res), p.two_pi);
        if (ry > p.pi) ry = __fadd_rn(ry, -p.two_pi);
    }
    off += 2 * p.nhb;
    const float h = __fadd_rn(__fmul_rn(reg[off], p.anchor[0]), p.anchor[0]);
    const float w = __fadd_rn(__fmul_rn(reg[off + 1], p.anchor[1]), p.anchor[1]);
    const float l = __fadd_rn(__fmul_rn(reg[off + 2], p.anchor[2]), p.anchor[2]);
    if (p.roi_dim == 7) {
        const float roi_ry = roi[6];
        const float ang = -roi_ry;
        const float cosa = cosf(ang), sina = sinf(ang);
        const float x2 = dot2(pos_x, cosa, pos_z, -sina, p.rot_mode);
        const float z2 = dot2(pos_x, sina, pos_z, cosa, p.rot_mode);
        pos_x = x2;
        pos_z = z2;
        ry = __fadd_rn(ry, roi_ry);
    }
    pos_x = __fadd_rn(pos_x, roi[0]);
    pos_z = __fadd_rn(pos_z, roi[2]);
    if (p.y_bottom) pos_y = __fadd_rn(pos_y, __fmul_rn(h, 0.5f));      // h / 2 is exact either way
    o[0] = pos_x; o[1] = pos_y; o[2] = pos_z; o[3] = h; o[4] = w; o[5] = l; o[6] = ry;
}

constexpr int kDecRows = 128;

// 128 rows per CTA: the (rows, c) block is contiguous in memory -> coalesced copy into padded shared rows, then one
// thread per row (a thread reading its own 300-byte row from global memory would touch 32 lines per warp request)
__global__ void __launch_bounds__(kDecRows) decode_kernel(const DecodeParams p) {
    extern __shared__ float rows_s[];
    const int ldc = p.c + 1;
    const long long row0 = (long long)blockIdx.x * kDecRows;
    const int nrows = (int)min((long long)kDecRows, p.rows - row0);
    const float *src = p.reg + row0 * p.c;
    const int total = nrows * p.c;
    for (int e = threadIdx.x; e < total; e += kDecRows) {
        const int r = e / p.c, col = e - r * p.c;
        rows_s[r * ldc + col] = __ldg(src + e);
    }
    __syncthreads();
    if ((int)threadIdx.x < nrows) {
        const long long row = row0 + threadIdx.x;
        float roi[7];
        for (int i = 0; i < p.roi_dim; ++i) roi[i] = __ldg(p.roi + row * p.roi_dim + i);
        float o[7];
        decode_row(p, rows_s + threadIdx.x * ldc, roi, o);
        float *dst = p.out + row * 7;
#pragma unroll
        for (int i = 0; i < 7; ++i) dst[i] = o[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// distance_based_proposal (proposal_layer.py:58-100): walk the points of a scene in descending score order, split them
// into the (0, 40] and (40, 80] m bands on the decoded z, keep the first pre0 / pre1 of each; an empty far band borrows
// the near-band members that follow the near quota.  One CTA per scene; ranks by ballot + shared prefix, so the
// candidate order is the sorted order exactly as torch's masked indexing produces it.
// Outputs per band: cidx (B, pre) point index of the candidate, bev (B, pre, 5) its BEV box
// (kitti_utils.py:boxes3d_to_bev_torch), cnt (2, B) how many (band-major, so each band's counts are contiguous).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;

__global__ void __launch_bounds__(kSelThreads) proposal_select_kernel(const long long *__restrict__ order,
                                                                      const float *__restrict__ props, int n, int pre0,
                                                                      int pre1, int32_t *__restrict__ cidx0,
                                                                      int32_t *__restrict__ cidx1, float *__restrict__ bev0,
                                                                      float *__restrict__ bev1, int32_t *__restrict__ cnt) {
    __shared__ int wsum[2][kSelThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    order += (size_t)b * n;
    props += (size_t)b * n * 7;
    cidx0 += (size_t)b * pre0; cidx1 += (size_t)b * pre1;
    bev0 += (size_t)b * pre0 * 5; bev1 += (size_t)b * pre1 * 5;

    // pass 1: is the far band empty?
    int far_any = 0;
    for (int i = tid; i < n; i += kSelThreads) {
        const float z = __ldg(props + (size_t)__ldg(order + i) * 7 + 2);
        far_any |= (z > 40.0f) && (z <= 80.0f);
    }
    const bool borrow = __syncthreads_or(far_any) == 0;

    int base_near = 0, base_far = 0;   // uniform
    for (int i0 = 0; i0 < n; i0 += kSelThreads) {
        const int i = i0 + tid;
        int j = 0;
        float z = 0.f;
        if (i < n) {
            j = (int)__ldg(order + i);
            z = __ldg(props + (size_t)j * 7 + 2);
        }
        const bool near = i < n && (z > 0.f) && (z <= 40.0f);
        const bool far = i < n && (z > 40.0f) && (z <= 80.0f);
        const unsigned bn = __ballot_sync(0xffffffffu, near), bf = __ballot_sync(0xffffffffu, far);
        if (lane == 0) { wsum[0][warp] = __popc(bn); wsum[1][warp] = __popc(bf); }
        __syncthreads();
        int on = base_near, of = base_far, tn = 0, tf = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) {
            const int cn = wsum[0][w], cf = wsum[1][w];
            if (w < warp) { on += cn; of += cf; }
            tn += cn; tf += cf;
        }
        const unsigned below = (1u << lane) - 1u;
        const int near_rank = on + __popc(bn & below), far_rank = of + __popc(bf & below);
        int band = -1, slot = 0;
        if (near) {
            if (near_rank < pre0) { band = 0; slot = near_rank; }
            else if (borrow && near_rank - pre0 < pre1) { band = 1; slot = near_rank - pre0; }
        } else if (far && far_rank < pre1) {
            band = 1; slot = far_rank;
        }
        if (band >= 0) {
            const float *q = props + (size_t)j * 7;
            const float cu = __ldg(q), cv = __ldg(q + 2), hw = __fmul_rn(__ldg(q + 4), 0.5f), hl = __fmul_rn(__ldg(q + 5), 0.5f);
            float *bv = (band ? bev1 : bev0) + (size_t)slot * 5;
            bv[0] = __fadd_rn(cu, -hl); bv[1] = __fadd_rn(cv, -hw); bv[2] = __fadd_rn(cu, hl); bv[3] = __fadd_rn(cv, hw);
            bv[4] = __ldg(q + 6);
            (band ? cidx1 : cidx0)[slot] = j;
        }
        base_near += tn;
        base_far += tf;
        __syncthreads();
    }
    if (tid == 0) {
        cnt[b] = min(base_near, pre0);
        cnt[gridDim.x + b] = borrow ? max(0, min(base_near - pre0, pre1)) : min(base_far, pre1);
    }
}

// keep lists of the two bands -> ret_bbox3d (B, post_tot, 7), ret_scores (B, post_tot): band 0's first min(num0, post0)
// boxes, then band 1's first min(num1, post1), zero-padded (proposal_layer.py:107-119, :38-44)
__global__ void __launch_bounds__(128) proposal_assemble_kernel(const float *__restrict__ props, const float *__restrict__ scores,
                                                                int n, const int32_t *__restrict__ cidx0,
                                                                const int32_t *__restrict__ cidx1, int pre0, int pre1,
                                                                const long long *__restrict__ keep0,
                                                                const long long *__restrict__ keep1,
                                                                const int32_t *__restrict__ num0, const int32_t *__restrict__ num1,
                                                                int post0, int post1, float *__restrict__ rois,
                                                                float *__restrict__ roi_scores) {
    const int b = blockIdx.x;
    const int k0 = min(__ldg(num0 + b), post0), k1 = min(__ldg(num1 + b), post1);
    const int post_tot = post0 + post1;
    for (int t = threadIdx.x; t < post_tot; t += blockDim.x) {
        float box[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float s = 0.f;
        int j = -1;
        if (t < k0) j = __ldg(cidx0 + (size_t)b * pre0 + (int)__ldg(keep0 + (size_t)b * post0 + t));
        else if (t - k0 < k1) j = __ldg(cidx1 + (size_t)b * pre1 + (int)__ldg(keep1 + (size_t)b * post1 + (t - k0)));
        if (j >= 0) {
            const float *q = props + ((size_t)b * n + j) * 7;
#pragma unroll
            for (int i = 0; i < 7; ++i) box[i] = __ldg(q + i);
            s = __ldg(scores + (size_t)b * n + j);
        }
        float *dst = rois + ((size_t)b * post_tot + t) * 7;
#pragma unroll
        for (int i = 0; i < 7; ++i) dst[i] = box[i];
        roi_scores[(size_t)b * post_tot + t] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// eval_rcnn.py:516-535, 611-620 for one scene per CTA: decode the M rois (M <= 256), select sigmoid(raw) > thresh,
// order the selection by descending raw score (stable, like torch.sort(stable=True) on the -inf-masked key) ->
// boxes_sorted (B, M, 7), scores_sorted (B, M), bev (B, M, 5), counts (B).  sigmoid as ATen computes it: 1 / (1 + exp(-x)).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPostMax = 256;

__global__ void __launch_bounds__(kPostMax) rcnn_post_prepare_kernel(DecodeParams p, const float *__restrict__ cls, int m,
                                                                     float thresh, float *__restrict__ boxes_sorted,
                                                                     float *__restrict__ scores_sorted, float *__restrict__ bev,
                                                                     int32_t *__restrict__ counts) {
    extern __shared__ float sm[];
    float *rows_s = sm;                                  // m x (c + 1)
    float *key_s = sm + (size_t)m * (p.c + 1);           // m
    const int b = blockIdx.x, t = threadIdx.x;
    const int ldc = p.c + 1;
    const float *src = p.reg + (size_t)b * m * p.c;
    for (int e = t; e < m * p.c; e += blockDim.x) {
        const int r = e / p.c, col = e - r * p.c;
        rows_s[r * ldc + col] = __ldg(src + e);
    }
    float raw = 0.f, key = -FLT_MAX;
    bool sel = false;
    if (t < m) {
        raw = __ldg(cls + (size_t)b * m + t);
        const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw)));
        sel = sg > thresh;
        key = sel ? raw : -INFINITY;
        key_s[t] = key;
    }
    __syncthreads();
    const int nsel = __syncthreads_count(sel);
    if (t < m) {
        float roi[7], o[7];
        const float *rp = p.roi + ((size_t)b * m + t) * 7;
#pragma unroll
        for (int i = 0; i < 7; ++i) roi[i] = __ldg(rp + i);
        decode_row(p, rows_s + t * ldc, roi, o);
        // stable descending rank of this roi's key
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            const float kj = key_s[j];
            rank += (kj > key) || (kj == key && j < t);
        }
        float *bd = boxes_sorted + ((size_t)b * m + rank) * 7;
#pragma unroll
        for (int i = 0; i < 7; ++i) bd[i] = o[i];
        scores_sorted[(size_t)b * m + rank] = raw;
        const float hw = __fmul_rn(o[4], 0.5f), hl = __fmul_rn(o[5], 0.5f);
        float *bv = bev + ((size_t)b * m + rank) * 5;
        bv[0] = __fadd_rn(o[0], -hl); bv[1] = __fadd_rn(o[2], -hw); bv[2] = __fadd_rn(o[0], hl); bv[3] = __fadd_rn(o[2], hw);
        bv[4] = o[6];
    }
    if (t == 0) counts[b] = nsel;
}

// keep (B, M) int64 / num (B) from the NMS -> rec (B, M, 8) [box7, raw score], rows >= num zero
__global__ void __launch_bounds__(128) rcnn_post_assemble_kernel(const float *__restrict__ boxes_sorted,
                                                                 const float *__restrict__ scores_sorted,
                                                                 const long long *__restrict__ keep,
                                                                 const int32_t *__restrict__ num, int m, float *__restrict__ rec) {
    const int b = blockIdx.x;
    const int k = __ldg(num + b);
    for (int t = threadIdx.x; t < m; t += blockDim.x) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t < k) {
            const int j = (int)__ldg(keep + (size_t)b * m + t);
            const float *q = boxes_sorted + ((size_t)b * m + j) * 7;
#pragma unroll
            for (int i = 0; i < 7; ++i) v[i] = __ldg(q + i);
            v[7] = __ldg(scores_sorted + (size_t)b * m + j);
        }
        float *dst = rec + ((size_t)b * m + t) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = v[i];
    }
}

int fill_decode(DecodeParams &p, const float *roi, int roi_dim, const float *reg, int c, long long rows, double loc_scope,
                double loc_bin_size, int num_head_bin, const float *anchor_h, int get_xz_fine, int get_y_by_bin,
                double loc_y_scope, double loc_y_bin_size, int get_ry_fine, int y_bottom, int rot_mode, const char *who) {
    const double kPi = 3.141592653589793;     // numpy's np.pi
    if (!roi || !reg || !anchor_h || rows < 0 || (roi_dim != 3 && roi_dim != 7) || num_head_bin <= 0 || loc_bin_size <= 0 ||
        (get_y_by_bin && loc_y_bin_size <= 0)) {
        pn2_set_last_error(who);
        return PN2_ERR_INVALID;
    }
    p.roi = roi; p.roi_dim = roi_dim; p.reg = reg; p.c = c; p.rows = rows;
    p.nb = (int)(loc_scope / loc_bin_size) * 2;
    p.lbs = (float)loc_bin_size; p.lbs_half = (float)(loc_bin_size / 2); p.loc_scope = (float)loc_scope;
    p.xz_fine = get_xz_fine;
    p.y_by_bin = get_y_by_bin;
    p.nby = get_y_by_bin ? (int)(loc_y_scope / loc_y_bin_size) * 2 : 0;
    p.ybs = (float)loc_y_bin_size; p.ybs_half = (float)(loc_y_bin_size / 2); p.y_scope = (float)loc_y_scope;
    p.nhb = num_head_bin; p.ry_fine = get_ry_fine;
    const double apc = get_ry_fine ? (kPi / 2) / num_head_bin : (2 * kPi) / num_head_bin;
    p.apc = (float)apc; p.apc_half = (float)(apc / 2);
    p.quarter_pi = (float)(kPi / 4); p.two_pi = (float)(2 * kPi); p.pi = (float)kPi;
    for (int i = 0; i < 3; ++i) p.anchor[i] = anchor_h[i];
    p.y_bottom = y_bottom; p.rot_mode = rot_mode;
    const int expect = p.nb * (get_xz_fine ? 4 : 2) + (get_y_by_bin ? 2 * p.nby : 1) + 2 * num_head_bin + 3;
    if (expect != c) {
        pn2_set_last_error("decode: the regression vector width does not match the bin configuration");
        return PN2_ERR_INVALID;
    }
    return PN2_OK;
}

}  // namespace

// decode_bbox_target (bbox_transform.py:24-121).  roi (rows, roi_dim) f32: point xyz (roi_dim 3) or rois (7);
// reg (rows, c); h_anchor: HOST pointer to the 3 mean sizes; out (rows, 7).  y_bottom adds h / 2 to y afterwards
// (proposal_layer.py:23).  rot_mode selects the rounding of the K = 2 matmul in rotate_pc_along_y_torch (roi_dim 7 only).
PN2_API int pn2_decode_bbox_f32(const float *roi, int roi_dim, const float *reg, int c, float *out, long long rows,
                                double loc_scope, double loc_bin_size, int num_head_bin, const float *h_anchor,
                                int get_xz_fine, int get_y_by_bin, double loc_y_scope, double loc_y_bin_size,
                                int get_ry_fine, int y_bottom, int rot_mode, cudaStream_t stream) {
    DecodeParams p = {};
    const int rc = fill_decode(p, roi, roi_dim, reg, c, rows, loc_scope, loc_bin_size, num_head_bin, h_anchor, get_xz_fine,
                               get_y_by_bin, loc_y_scope, loc_y_bin_size, get_ry_fine, y_bottom, rot_mode,
                               "pn2_decode_bbox_f32: bad argument");
    if (rc) return rc;
    if (!out) { pn2_set_last_error("pn2_decode_bbox_f32: bad argument"); return PN2_ERR_INVALID; }
    if (rows == 0) return PN2_OK;
    p.out = out;
    const size_t smem = (size_t)kDecRows * (c + 1) * sizeof(float);
    if (smem > 48 * 1024) { pn2_set_last_error("pn2_decode_bbox_f32: regression vector too wide"); return PN2_ERR_UNSUPPORTED; }
    decode_kernel<<<(unsigned)((rows + kDecRows - 1) / kDecRows), kDecRows, smem, stream>>>(p);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// order (B, N) int64: the points of each scene by descending score (torch.sort); props (B, N, 7) decoded proposals.
// -> cidx0 (B, pre0) / cidx1 (B, pre1) int32, bev0 (B, pre0, 5) / bev1 (B, pre1, 5), cnt (2, B) int32 {near row, far row}.
PN2_API int pn2_proposal_select_f32(const long long *order, const float *props, int b, int n, int pre0, int pre1,
                                    int32_t *cidx0, int32_t *cidx1, float *bev0, float *bev1, int32_t *cnt,
                                    cudaStream_t stream) {
    if (!order || !props || !cidx0 || !cidx1 || !bev0 || !bev1 || !cnt || b < 0 || n <= 0 || pre0 <= 0 || pre1 <= 0) {
        pn2_set_last_error("pn2_proposal_select_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    proposal_select_kernel<<<b, kSelThreads, 0, stream>>>(order, props, n, pre0, pre1, cidx0, cidx1, bev0, bev1, cnt);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_proposal_assemble_f32(const float *props, const float *scores, int b, int n, const int32_t *cidx0,
                                      const int32_t *cidx1, int pre0, int pre1, const long long *keep0,
                                      const long long *keep1, const int32_t *num0, const int32_t *num1, int post0,
                                      int post1, float *rois, float *roi_scores, cudaStream_t stream) {
    if (!props || !scores || !cidx0 || !cidx1 || !keep0 || !keep1 || !num0 || !num1 || !rois || !roi_scores || b < 0 ||
        post0 < 0 || post1 < 0) {
        pn2_set_last_error("pn2_proposal_assemble_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    proposal_assemble_kernel<<<b, 128, 0, stream>>>(props, scores, n, cidx0, cidx1, pre0, pre1, keep0, keep1, num0, num1,
                                                   post0, post1, rois, roi_scores);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// rois (B, M, 7), reg (B * M, c), cls (B * M) raw scores of the single foreground class -> boxes_sorted (B, M, 7),
// scores_sorted (B, M), bev (B, M, 5) in descending-score order of the selected rois, counts (B) = how many are selected.
PN2_API int pn2_rcnn_post_prepare_f32(const float *rois, const float *reg, int c, const float *cls, int b, int m,
                                      double loc_scope, double loc_bin_size, int num_head_bin, const float *h_anchor,
                                      int get_y_by_bin, double loc_y_scope, double loc_y_bin_size, double score_thresh,
                                      int rot_mode, float *boxes_sorted, float *scores_sorted, float *bev, int32_t *counts,
                                      cudaStream_t stream) {
    DecodeParams p = {};
    const int rc = fill_decode(p, rois, 7, reg, c, (long long)b * m, loc_scope, loc_bin_size, num_head_bin, h_anchor, 1,
                               get_y_by_bin, loc_y_scope, loc_y_bin_size, 1, 0, rot_mode,
                               "pn2_rcnn_post_prepare_f32: bad argument");
    if (rc) return rc;
    if (!cls || !boxes_sorted || !scores_sorted || !bev || !counts || m <= 0 || m > kPostMax) {
        pn2_set_last_error("pn2_rcnn_post_prepare_f32: bad argument (at most 256 rois per scene)");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    const size_t smem = ((size_t)m * (c + 1) + m) * sizeof(float);
    if (smem > 48 * 1024) { pn2_set_last_error("pn2_rcnn_post_prepare_f32: too much shared memory"); return PN2_ERR_UNSUPPORTED; }
    rcnn_post_prepare_kernel<<<b, kPostMax, smem, stream>>>(p, cls, m, (float)score_thresh, boxes_sorted, scores_sorted, bev,
                                                           counts);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_rcnn_post_assemble_f32(const float *boxes_sorted, const float *scores_sorted, const long long *keep,
                                       const int32_t *num, int b, int m, float *rec, cudaStream_t stream) {
    if (!boxes_sorted || !scores_sorted || !keep || !num || !rec || b < 0 || m <= 0) {
        pn2_set_last_error("pn2_rcnn_post_assemble_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    rcnn_post_assemble_kernel<<<b, 128, 0, stream>>>(boxes_sorted, scores_sorted, keep, num, m, rec);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Descending argsort of the per-scene RPN scores (lib/rpn/proposal_layer.py:26: torch.sort(scores, dim=1, descending=True)).
// torch runs it as a segmented radix sort, eleven launches; here one CTA sorts one row in shared memory: bitonic network
// over 64-bit keys (~sortable score bits | index), i.e. descending score, ascending index among equal scores -- the stable
// order the radix sort produces (-0 and +0 are one key, as in its key transform).  N <= 16384.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kSortThreads = 512;

// position of key i in shared memory: one padding word per 16 keys, so that the 16-key register blocks of a warp's lanes
// (128 bytes apart without it) fall into different banks
__device__ __forceinline__ int sort_pos(int i) { return i + (i >> 4); }

// Substeps of one bitonic stage `size` with strides 2^(lo + NSUB - 1) .. 2^lo, in registers: a thread loads the 2^NSUB keys
// whose indices differ only in bits [lo, lo + NSUB), runs the NSUB compare-exchange substeps and stores them back --
// one shared-memory round trip for up to four substeps (the plain network is bound by shared-memory bandwidth:
// 105 round trips of 128 KB for 16384 keys; this way 32).
template <int NSUB>
__device__ __forceinline__ void sort_block_steps(unsigned long long *keys, int npow2, int size, int lo, int tid) {
    constexpr int E = 1 << NSUB;
    const int nblocks = npow2 >> NSUB;
    for (int blk = tid; blk < nblocks; blk += kSortThreads) {
        const int low = blk & ((1 << lo) - 1), high = blk >> lo;
        const int base = (high << (lo + NSUB)) | low;
        const bool up = (base & size) == 0;
        unsigned long long r[E];
#pragma unroll
        for (int j = 0; j < E; ++j) r[j] = keys[sort_pos(base + (j << lo))];
#pragma unroll
        for (int bit = NSUB - 1; bit >= 0; --bit) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if ((j >> bit) & 1) continue;
                const unsigned long long a = r[j], c = r[j | (1 << bit)];
                const bool sw = (a > c) == up;
                r[j] = sw ? c : a;
                r[j | (1 << bit)] = sw ? a : c;
            }
        }
#pragma unroll
        for (int j = 0; j < E; ++j) keys[sort_pos(base + (j << lo))] = r[j];
    }
}

__global__ void __launch_bounds__(kSortThreads) argsort_desc_kernel(const float *__restrict__ scores, long long *__restrict__ order,
                                                                   int n, int npow2) {
    extern __shared__ unsigned long long keys[];
    const float *row = scores + (size_t)blockIdx.x * n;
    long long *out = order + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < npow2; i += kSortThreads) {
        unsigned long long k = 0xFFFFFFFFFFFFFFFFull;             // padding sorts last
        if (i < n) {
            uint32_t u = __float_as_uint(__ldg(row + i));
            if (u == 0x80000000u) u = 0u;                         // -0 == +0, as for the radix sort's key transform
            const uint32_t asc = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);   // monotone in the float order
            k = ((unsigned long long)(~asc) << 32) | (uint32_t)i;
        }
        keys[sort_pos(i)] = k;
    }
    __syncthreads();
    int s = 1;
    for (int size = 2; size <= npow2; size <<= 1, ++s) {
        // stride exponents s-1 .. 0: a first group of ((s - 1) % 4) + 1 substeps, then groups of four
        int hi = s;                                   // exponents [hi - g, hi) are done next
        const int g0 = ((s - 1) & 3) + 1;
        switch (g0) {
            case 1: sort_block_steps<1>(keys, npow2, size, hi - 1, threadIdx.x); break;
            case 2: sort_block_steps<2>(keys, npow2, size, hi - 2, threadIdx.x); break;
            case 3: sort_block_steps<3>(keys, npow2, size, hi - 3, threadIdx.x); break;
            default: sort_block_steps<4>(keys, npow2, size, hi - 4, threadIdx.x); break;
        }
        hi -= g0;
        __syncthreads();
        for (; hi > 0; hi -= 4) {
            sort_block_steps<4>(keys, npow2, size, hi - 4, threadIdx.x);
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += kSortThreads) out[i] = (long long)(uint32_t)keys[sort_pos(i)];
}
}  // namespace

// scores (B, N) f32 -> order (B, N) int64: indices by descending score, ties by ascending index.  N <= 16384.
PN2_API int pn2_argsort_desc_f32(const float *scores, long long *order, int b, int n, cudaStream_t stream) {
    if (b < 0 || n < 0 || (b * (long long)n > 0 && (!scores || !order))) {
        pn2_set_last_error("pn2_argsort_desc_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (n > 16384) {
        pn2_set_last_error("pn2_argsort_desc_f32: rows of more than 16384 scores are not supported");
        return PN2_ERR_UNSUPPORTED;
    }
    if (b == 0 || n == 0) return PN2_OK;
    int npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    const size_t smem = (size_t)(npow2 + (npow2 >> 4) + 1) * sizeof(unsigned long long);
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(argsort_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (16384 + 1024 + 1) * 8) != cudaSuccess) {
            pn2_set_last_error("pn2_argsort_desc_f32: cudaFuncSetAttribute failed");
            return PN2_ERR_LAUNCH;
        }
        attr_done = true;
    }
    argsort_desc_kernel<<<b, kSortThreads, smem, stream>>>(scores, order, n, npow2);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
