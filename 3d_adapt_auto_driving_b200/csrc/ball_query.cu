// ball_query.cu -- radius neighbour search for sm_100a.
//
// Replaces pointrcnn/pointnet2_lib/pointnet2/src/ball_query_gpu.cu:9-67 behind
// pn2_ball_query_f32 (one radius, the reference API) and pn2_ball_query_dual_f32 (both MSG
// scales of one SA layer in a single scan; the reference runs two launches).
//
// Semantics kept bit-for-bit: candidates are visited in ascending index, a point is taken iff
// d2 < radius*radius (strict, f32, the reference's FMA order), the first hit fills every slot
// of the row, later hits overwrite slots 1.., rows without a hit are NOT written (the caller
// zero-initialises, pointnet2_utils.py:218), the scan stops at nsample hits.
//
// Design: one thread per query centre, the cloud streamed through shared memory in tiles as
// float4 (one conflict-free broadcast LDS.128 per candidate).  The reference spends 3 global
// loads + 6 FP ops on every (centre, point) pair; here a pair first takes a one-subtract /
// one-compare rejection on |dx| >= r.  That test is exact: rounding is monotone, so
// fl(dx*dx + t) >= fl(r*r) whenever |dx| >= |r| and t >= 0, and the later fma only adds a
// non-negative term; the full distance is evaluated only for the ~2r/extent fraction that
// survives.  A CTA leaves the tile loop as soon as all its centres are full.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;   // centres per CTA
constexpr int kTile = 1024;     // candidates per shared-memory tile (16 KB)

template <bool DUAL>
__global__ void __launch_bounds__(kThreads) ball_query_kernel(const float *__restrict__ new_xyz,
                                                             const float *__restrict__ xyz, int32_t *__restrict__ idx0,
                                                             int32_t *__restrict__ idx1, int n, int m, float radius0,
                                                             int ns0, float radius1, int ns1) {
    __shared__ float4 tile[kTile];
    const int cloud = blockIdx.y;
    const int c = blockIdx.x * kThreads + threadIdx.x;
    const bool active = c < m;
    xyz += (size_t)cloud * n * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *q = new_xyz + ((size_t)cloud * m + c) * 3;
        qx = __ldg(q + 0); qy = __ldg(q + 1); qz = __ldg(q + 2);
    }
    const float r2_0 = __fmul_rn(radius0, radius0);
    const float r2_1 = DUAL ? __fmul_rn(radius1, radius1) : 0.f;
    const float rmax = DUAL ? fmaxf(fabsf(radius0), fabsf(radius1)) : fabsf(radius0);
    int32_t *row0 = idx0 + ((size_t)cloud * m + (active ? c : 0)) * ns0;
    int32_t *row1 = DUAL ? idx1 + ((size_t)cloud * m + (active ? c : 0)) * ns1 : nullptr;
    int cnt0 = active ? 0 : ns0;
    int cnt1 = (DUAL && active) ? 0 : ns1;
    if (!DUAL) cnt1 = 0x7fffffff;

    for (int base = 0; base < n; base += kTile) {
        const int len = min(kTile, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += kThreads) {
            const float *p = xyz + (size_t)(base + i) * 3;
            tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        const bool done = (cnt0 >= ns0) && (!DUAL || cnt1 >= ns1);
        if (__syncthreads_and(done)) break;
        if (done) continue;
#pragma unroll 4
        for (int i = 0; i < len; ++i) {
            const float4 p = tile[i];
            const float dx = qx - p.x;
            if (fabsf(dx) < rmax) {
                const float d2 = pn2_sqdist(dx, qy - p.y, qz - p.z);
                const int k = base + i;
                if (d2 < r2_0 && cnt0 < ns0) {
                    if (cnt0 == 0)
                        for (int l = 0; l < ns0; ++l) row0[l] = k;
                    else
                        row0[cnt0] = k;
                    ++cnt0;
                }
                if (DUAL && d2 < r2_1 && cnt1 < ns1) {
                    if (cnt1 == 0)
                        for (int l = 0; l < ns1; ++l) row1[l] = k;
                    else
                        row1[cnt1] = k;
                    ++cnt1;
                }
            }
        }
    }
}

}  // namespace

PN2_API int pn2_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n, int m, float radius,
                               int nsample, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0) {
        pn2_set_last_error("pn2_ball_query_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return PN2_OK;
    dim3 grid(pn2_divup(m, kThreads), b);
    ball_query_kernel<false><<<grid, kThreads, 0, stream>>>(new_xyz, xyz, idx, nullptr, n, m, radius, nsample, 0.f, 0);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_ball_query_dual_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int b, int n,
                                    int m, float radius0, int nsample0, float radius1, int nsample1,
                                    cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample0 <= 0 || nsample1 <= 0) {
        pn2_set_last_error("pn2_ball_query_dual_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0 || n == 0) return PN2_OK;
    dim3 grid(pn2_divup(m, kThreads), b);
    ball_query_kernel<true><<<grid, kThreads, 0, stream>>>(new_xyz, xyz, idx0, idx1, n, m, radius0, nsample0, radius1,
                                                           nsample1);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
