// roipool3d.cu -- point-in-ROI pooling for sm_100a.
//
// Replaces pointrcnn/lib/utils/roipool3d/src/roipool3d_kernel.cu:97-232 (assign_pts_to_box3d +
// get_pooled_idx + roipool3d_forward, plus the launcher's per-call cudaMalloc of a B*N*M int
// flag tensor, 105 MB at B=16) behind pn2_roipool3d_f32.
//
// One CTA per ROI.  The CTA walks the cloud in index order, 256 points per step; each warp
// compacts its hits with a ballot, the warps' counts are prefix-summed through shared memory,
// and the first `sampled` hit indices land in a shared list -- the same list the reference's
// one-thread-per-ROI scan builds -- after which the scan stops.  The list is wrapped
// (k % cnt, roipool3d_kernel.cu:152-158) and the rows [xyz | features] are copied out with one
// warp per sampled point so both the gather read and the pooled write are coalesced.
// No flag tensor, no scratch, one launch.
//
// The in-box predicate reproduces the arithmetic of the compiled reference exactly
// (roipool3d_kernel.cu:14-28 as it comes out of nvcc 12.9): half extents and the box centre
// height are formed in double from the float inputs, the rotated offsets are
//   x_rot = fma(dx, cos, -fl(dz*sin)),  z_rot = fma(dz, cos, fl(dx*sin))
// with cosf/sinf of the box angle, and the final four compares are done in double.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct BoxTest {
    float cx, cz, cy;     // centre (cy already lowered by h/2, rounded to float like the reference)
    double hh, hl, hw;    // half extents in double
    float cosa, sina;
};

__device__ __forceinline__ bool in_box(const BoxTest &b, float x, float y, float z) {
    const float dx = x - b.cx;
    if (fabsf(dx) > 10.0f) return false;
    if (b.hh < (double)fabsf(y - b.cy)) return false;
    const float dz = z - b.cz;
    if (fabsf(dz) > 10.0f) return false;
    const float x_rot = __fmaf_rn(dx, b.cosa, -__fmul_rn(dz, b.sina));
    const float z_rot = __fmaf_rn(dz, b.cosa, __fmul_rn(dx, b.sina));
    return ((double)x_rot >= -b.hl) & ((double)x_rot <= b.hl) & ((double)z_rot >= -b.hw) & ((double)z_rot <= b.hw);
}

__global__ void __launch_bounds__(kThreads) roipool3d_kernel(const float *__restrict__ xyz,
                                                            const float *__restrict__ boxes3d,
                                                            const float *__restrict__ feat,
                                                            const float *__restrict__ feat2,
                                                            float *__restrict__ pooled, int32_t *__restrict__ empty,
                                                            int n, int m, int c, int c2, int off2, int row, int sampled) {
    extern __shared__ int32_t list[];  // sampled entries
    __shared__ int warp_cnt[kWarps];

    const int roi = blockIdx.x, cloud = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *bx = boxes3d + ((size_t)cloud * m + roi) * 7;
    xyz += (size_t)cloud * n * 3;
    feat += (size_t)cloud * n * c;
    if (feat2) feat2 += (size_t)cloud * n * c2;

    BoxTest b;
    {
        const float by = __ldg(bx + 1), h = __ldg(bx + 3), w = __ldg(bx + 4), l = __ldg(bx + 5), ang = __ldg(bx + 6);
        b.cx = __ldg(bx + 0);
        b.cz = __ldg(bx + 2);
        b.hh = (double)h * 0.5;
        b.hl = (double)l * 0.5;
        b.hw = (double)w * 0.5;
        b.cy = (float)((double)by - b.hh);
        b.cosa = cosf(ang);
        b.sina = sinf(ang);
    }
    __syncthreads();

    int cnt = 0;  // uniform across the CTA
    for (int base = 0; base < n && cnt < sampled; base += kThreads) {
        const int k = base + tid;
        bool hit = false;
        if (k < n) hit = in_box(b, __ldg(xyz + (size_t)k * 3), __ldg(xyz + (size_t)k * 3 + 1), __ldg(xyz + (size_t)k * 3 + 2));
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = cnt;
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) {
            const int cw = warp_cnt[wv];
            if (wv < warp) off += cw;
            cnt += cw;
        }
        if (hit) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            if (pos < sampled) list[pos] = k;
        }
        __syncthreads();
    }
    if (cnt == 0) {  // the (pre-zeroed) pooled rows stay zero, only the flag is raised
        if (tid == 0) empty[(size_t)cloud * m + roi] = 1;
        return;
    }
    const int have = min(cnt, sampled);
    float *dst_base = pooled + ((size_t)cloud * m + roi) * (size_t)sampled * row;
    for (int s = warp; s < sampled; s += kWarps) {
        const int src = list[s < have ? s : s % have];
        float *dst = dst_base + (size_t)s * row;
        const float *f = feat + (size_t)src * c;
        if (lane < 3) dst[lane] = __ldg(xyz + (size_t)src * 3 + lane);
        for (int j = lane; j < c; j += 32) dst[3 + j] = __ldg(f + j);
        if (feat2) {   // second feature block at a 16-byte aligned column: 128-bit copies
            const float4 *f2 = reinterpret_cast<const float4 *>(feat2 + (size_t)src * c2);
            float4 *d2 = reinterpret_cast<float4 *>(dst + off2);
            for (int j = lane; j < (c2 >> 2); j += 32) d2[j] = __ldg(f2 + j);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The RCNN input stage in ONE launch (rcnn_net.py:126-154 with cfg USE_MASK, USE_DEPTH, no intensity):
//   enlarge_box3d (kitti_utils.py:150-160) -> point-in-box pooling of 512 points per ROI -> canonical transform
//   (pooled xyz -= roi centre; rotate_pc_along_y_torch by the roi angle, kitti_utils.py:45-63),
// with the per-point extras [seg mask = sigmoid(score) > thresh, depth / 70 - 0.5] formed on the fly from the raw RPN
// score and the point norm, so neither the (B, N, 2 + C) feature concat, nor the zero-fill of the 0.44 GB pooled tensor,
// nor the read-modify-write passes of the transform exist.  Every row of every ROI is written (an empty ROI gives the
// transformed zero rows the reference produces); the row layout is pn2_roipool3d_split_f32's
//   [x y z | mask depth | 0-pad to off2 | feat2 (c2)].
// Float semantics are torch's: the scalar divisor becomes a multiplication by fl(1/70) (ATen div by a CPU scalar), the
// K = 2 batched matmul of the rotation rounds as fma(z, r1, fl(x * r0)) for this shape (rot_mode 0; 1 / 2: the other
// orders, see glue.py) -- pinned against torch on the GPU by tests/test_glue_gpu.py.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rp_dot2(float a0, float b0, float a1, float b1, int mode) {
    if (mode == 0) return __fmaf_rn(a1, b1, __fmul_rn(a0, b0));
    if (mode == 1) return __fmaf_rn(a0, b0, __fmul_rn(a1, b1));
    return __fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}

__global__ void __launch_bounds__(kThreads) roipool3d_canon_kernel(const float *__restrict__ xyz, const float *__restrict__ rois,
                                                                  const float *__restrict__ score, const float *__restrict__ depth,
                                                                  const float *__restrict__ feat2, float *__restrict__ pooled,
                                                                  int32_t *__restrict__ empty, int n, int m, int c2, int off2,
                                                                  int row, int sampled, float extra, float extra2, float thresh,
                                                                  float inv_depth, int rot_mode) {
    extern __shared__ int32_t list[];          // sampled hit list | sampled source ids | sampled x 5 head values
    __shared__ int warp_cnt[4 * kWarps];
    const int roi = blockIdx.x, cloud = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *bx = rois + ((size_t)cloud * m + roi) * 7;
    xyz += (size_t)cloud * n * 3;
    score += (size_t)cloud * n;
    depth += (size_t)cloud * n;
    feat2 += (size_t)cloud * n * c2;

    const float rx = __ldg(bx + 0), ry = __ldg(bx + 1), rz = __ldg(bx + 2), ang = __ldg(bx + 6);
    BoxTest b;
    {
        // enlarge_box3d: h, w, l += 2 * extra; y += extra (float adds of the rounded Python scalars)
        const float by = __fadd_rn(ry, extra), h = __fadd_rn(__ldg(bx + 3), extra2), w = __fadd_rn(__ldg(bx + 4), extra2),
                    l = __fadd_rn(__ldg(bx + 5), extra2);
        b.cx = rx;
        b.cz = rz;
        b.hh = (double)h * 0.5;
        b.hl = (double)l * 0.5;
        b.hw = (double)w * 0.5;
        b.cy = (float)((double)by - b.hh);
        b.cosa = cosf(ang);
        b.sina = sinf(ang);
    }
    __syncthreads();

    // ---- phase 1: ordered list of the first `sampled` points inside the box.  kScan points per thread and step, all loads
    //      issued before the tests (the box test's early exits are shortcuts, evaluating every condition gives the same
    //      boolean): a step costs one load latency and two barriers for 1024 points.  (First version: one point per thread
    //      with the x / y / z loads serialised behind the early exits -- 64 steps of three dependent L2 round trips.) ----
    constexpr int kScan = 4;
    int32_t *srcs = list + sampled;                                      // phase 2: source point of every pooled row
    float *head = reinterpret_cast<float *>(list + 2 * sampled);         // phase 2: 5 head values per pooled row
    int cnt = 0;
    for (int base = 0; base < n && cnt < sampled; base += kThreads * kScan) {
        float px[kScan], py[kScan], pz[kScan];
#pragma unroll
        for (int u = 0; u < kScan; ++u) {
            const int k = min(base + u * kThreads + tid, n - 1);
            px[u] = __ldg(xyz + (size_t)k * 3); py[u] = __ldg(xyz + (size_t)k * 3 + 1); pz[u] = __ldg(xyz + (size_t)k * 3 + 2);
        }
        unsigned bal[kScan];
        int mine = 0;                       // lane u < kScan of every warp publishes the warp's hit count of sub-step u
#pragma unroll
        for (int u = 0; u < kScan; ++u) {
            const bool hit = base + u * kThreads + tid < n && in_box(b, px[u], py[u], pz[u]);
            bal[u] = __ballot_sync(0xffffffffu, hit);
            if (lane == u) mine = __popc(bal[u]);
        }
        if (lane < kScan) warp_cnt[lane * kWarps + warp] = mine;
        __syncthreads();
        // hits are ordered by point index: sub-step u (points base + u * 256 ..), then warp, then lane
        int off = cnt;
#pragma unroll
        for (int u = 0; u < kScan; ++u) {
            int before = 0, total = 0;
#pragma unroll
            for (int wv = 0; wv < kWarps; ++wv) {
                const int cw = warp_cnt[u * kWarps + wv];
                if (wv < warp) before += cw;
                total += cw;
            }
            if ((bal[u] >> lane) & 1u) {
                const int pos = off + before + __popc(bal[u] & ((1u << lane) - 1u));
                if (pos < sampled) list[pos] = base + u * kThreads + tid;
            }
            off += total;
        }
        cnt = off;
        __syncthreads();
    }
    if (cnt == 0 && tid == 0) empty[(size_t)cloud * m + roi] = 1;
    const int have = min(cnt, sampled);

    // ---- phase 2: the five head values of every pooled row, one THREAD per row (uniform code; in the first version every
    //      warp walked through the lane-0..4 branches of each of its rows: ~150 instructions a row) ----
    // rotation of the canonical transform: the same cos / sin of the roi angle (torch.cos / torch.sin of a float tensor)
    const float cosa = b.cosa, sina = b.sina, nsina = -sina;
    for (int s2 = tid; s2 < sampled; s2 += kThreads) {
        float px = 0.f, py = 0.f, pz = 0.f, mk = 0.f, dp = 0.f;
        int src = -1;
        if (have > 0) {
            src = list[s2 < have ? s2 : s2 % have];
            px = __ldg(xyz + (size_t)src * 3);
            py = __ldg(xyz + (size_t)src * 3 + 1);
            pz = __ldg(xyz + (size_t)src * 3 + 2);
            const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__ldg(score + src))));
            mk = sg > thresh ? 1.0f : 0.0f;
            dp = __fadd_rn(__fmul_rn(__ldg(depth + src), inv_depth), -0.5f);
        }
        const float x = __fadd_rn(px, -rx), z = __fadd_rn(pz, -rz);
        float *hd = head + s2 * 5;
        hd[0] = rp_dot2(x, cosa, z, nsina, rot_mode);
        hd[1] = __fadd_rn(py, -ry);
        hd[2] = rp_dot2(x, sina, z, cosa, rot_mode);
        hd[3] = mk;
        hd[4] = dp;
        srcs[s2] = src;
    }
    __syncthreads();

    // ---- phase 3: the rows, one warp per row, four rows in flight per warp ----
    float *dst_base = pooled + ((size_t)cloud * m + roi) * (size_t)sampled * row;
    const int nv = c2 >> 2;
    constexpr int kRows = 4;
    for (int s0 = warp * kRows; s0 < sampled; s0 += kWarps * kRows) {
        float4 v[kRows];
        float hv[kRows];
        int src[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            const int sr = min(s0 + u, sampled - 1);
            src[u] = srcs[sr];
            hv[u] = lane < 5 ? head[sr * 5 + lane] : 0.f;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src[u] >= 0 && lane < nv) v[u] = __ldg(reinterpret_cast<const float4 *>(feat2 + (size_t)src[u] * c2) + lane);
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            if (s0 + u >= sampled) break;
            float *dst = dst_base + (size_t)(s0 + u) * row;
            if (lane < off2) dst[lane] = hv[u];
            float4 *d2 = reinterpret_cast<float4 *>(dst + off2);
            if (lane < nv) d2[lane] = v[u];
            for (int j = lane + 32; j < nv; j += 32)          // feature blocks wider than 128 channels
                d2[j] = src[u] >= 0 ? __ldg(reinterpret_cast<const float4 *>(feat2 + (size_t)src[u] * c2) + j)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = off2 + c2 + lane; j < row; j += 32) dst[j] = 0.f;
        }
    }
}

}  // namespace

// xyz (B,N,3), boxes3d (B,M,7) ALREADY enlarged by the caller (roipool3d_utils.py:18),
// feat (B,N,C) -> pooled (B,M,sampled,3+C) and empty (B,M) int32, both pre-zeroed by the
// caller exactly as roipool3d_utils.py:20-22 does.
PN2_API int pn2_roipool3d_f32(const float *xyz, const float *boxes3d, const float *feat, float *pooled, int32_t *empty,
                              int b, int n, int m, int c, int sampled, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || c < 0 || sampled <= 0 || sampled > 8192) {
        pn2_set_last_error("pn2_roipool3d_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0) return PN2_OK;
    dim3 grid(m, b);
    roipool3d_kernel<<<grid, kThreads, sampled * sizeof(int32_t), stream>>>(xyz, boxes3d, feat, nullptr, pooled, empty, n,
                                                                           m, c, 0, 0, 3 + c, sampled);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// The same pooling with the per-point feature vector given in two pieces and a padded output
// row, so that the wide piece lands 16-byte aligned for the tensor-core MLP that consumes it:
//   pooled row (ld_out floats) = [x y z | feat (c) | zero pad ... | feat2 (c2) at column off2 | pad]
// feat (B,N,c) may be NULL when c == 0.  Needs c2 % 4 == 0, off2 % 4 == 0, off2 >= 3 + c,
// ld_out % 4 == 0, ld_out >= off2 + c2 and 16-byte aligned feat2 / pooled.  `pooled` pre-zeroed.
// Selection and order of the sampled points are those of pn2_roipool3d_f32.
PN2_API int pn2_roipool3d_split_f32(const float *xyz, const float *boxes3d, const float *feat, int c, const float *feat2,
                                    int c2, float *pooled, int ld_out, int off2, int32_t *empty, int b, int n, int m,
                                    int sampled, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || c < 0 || c2 <= 0 || !feat2 || (c > 0 && !feat) || sampled <= 0 || sampled > 8192 ||
        off2 < 3 + c || ld_out < off2 + c2) {
        pn2_set_last_error("pn2_roipool3d_split_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if ((c2 & 3) || (off2 & 3) || (ld_out & 3) || (reinterpret_cast<uintptr_t>(feat2) & 15) ||
        (reinterpret_cast<uintptr_t>(pooled) & 15)) {
        pn2_set_last_error("pn2_roipool3d_split_f32: feat2 block must be 16-byte aligned");
        return PN2_ERR_UNSUPPORTED;
    }
    if (b == 0 || m == 0) return PN2_OK;
    dim3 grid(m, b);
    roipool3d_kernel<<<grid, kThreads, sampled * sizeof(int32_t), stream>>>(xyz, boxes3d, feat, feat2, pooled, empty, n,
                                                                           m, c, c2, off2, ld_out, sampled);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// rcnn_net.py:126-154 in one launch (see roipool3d_canon_kernel): rois (B, M, 7) NOT enlarged, score (B, N) raw RPN
// foreground score, depth (B, N) point norm, feat2 (B, N, c2) point features -> pooled (B, M, sampled, ld_out) rows
// [canonical x y z | mask | depth / 70 - 0.5 | 0-pad | feat2 at column off2 | 0-pad], empty (B, M) int32 ZEROED by the
// caller.  `pooled` needs no pre-zeroing.  Needs off2 in [5, 32], c2 % 4 == 0, off2 % 4 == 0, ld_out % 4 == 0.
PN2_API int pn2_roipool3d_canon_f32(const float *xyz, const float *rois, double extra_width, const float *score,
                                    double score_thresh, const float *depth, double depth_norm, const float *feat2, int c2,
                                    float *pooled, int ld_out, int off2, int32_t *empty, int b, int n, int m, int sampled,
                                    int rot_mode, cudaStream_t stream) {
    if (!xyz || !rois || !score || !depth || !feat2 || !pooled || !empty || b < 0 || n < 0 || m < 0 || c2 <= 0 ||
        sampled <= 0 || sampled > 8192 || off2 < 5 || off2 > 32 || ld_out < off2 + c2 || depth_norm == 0.0) {
        pn2_set_last_error("pn2_roipool3d_canon_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if ((c2 & 3) || (off2 & 3) || (ld_out & 3) || (reinterpret_cast<uintptr_t>(feat2) & 15) ||
        (reinterpret_cast<uintptr_t>(pooled) & 15)) {
        pn2_set_last_error("pn2_roipool3d_canon_f32: feat2 block must be 16-byte aligned");
        return PN2_ERR_UNSUPPORTED;
    }
    if (sampled > 1024) {
        pn2_set_last_error("pn2_roipool3d_canon_f32: more than 1024 sampled points per roi are not supported (28 bytes of shared memory each)");
        return PN2_ERR_UNSUPPORTED;
    }
    if (b == 0 || m == 0) return PN2_OK;
    dim3 grid(m, b);
    // ATen divides by a CPU scalar as a multiplication by the reciprocal formed in float
    const float inv = 1.0f / (float)depth_norm;
    roipool3d_canon_kernel<<<grid, kThreads, (size_t)sampled * 7 * sizeof(int32_t), stream>>>(
        xyz, rois, score, depth, feat2, pooled, empty, n, m, c2, off2, ld_out, sampled, (float)extra_width,
        (float)(extra_width * 2), (float)score_thresh, inv, rot_mode);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
