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
