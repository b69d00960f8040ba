// interpolate.cu -- three nearest neighbours + 3-tap interpolation for sm_100a.
//
// Replaces pointrcnn/pointnet2_lib/pointnet2/src/interpolate_gpu.cu:9-160 behind
// pn2_three_nn_f32 / pn2_three_interpolate_f32 / pn2_three_interpolate_grad_f32.
//
// three_nn keeps the reference's result exactly: ascending scan, strict '<' against the three
// running bests (so the lowest index wins ties), outputs SQUARED distances.  The reference
// keeps the bests in double initialised to 1e40 and compares the promoted float distance;
// with finite inputs that is the same order as float bests initialised to +inf, and
// (float)1e40 == +inf, so the outputs are identical bit for bit.
// Design: one thread per unknown point, known points streamed through shared memory as float4;
// a candidate is rejected after one subtract + one multiply when dx*dx >= best3 (exact:
// fl(dx*dx + t) >= fl(dx*dx) for t >= 0), which removes the other 4 FP ops and 3 compares for
// almost every pair once best3 has tightened.
//
// three_interpolate: out[b,c,i] = fma(w2,p2, fma(w0,p0, w1*p1)), the contraction nvcc picked for
// the reference expression (interpolate_gpu.cu:96).  One thread per unknown point loops over
// a block of channels so idx/weight are read once per 8 channels instead of once per channel.
#include "common.cuh"
#include <cstdlib>
#include "spatial_order.cuh"
#include <math.h>

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 1024;

__global__ void __launch_bounds__(kThreads) three_nn_kernel(const float *__restrict__ unknown,
                                                           const float *__restrict__ known, float *__restrict__ dist2,
                                                           int32_t *__restrict__ idx, int n, int m) {
    __shared__ float4 tile[kTile];
    const int cloud = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const bool active = i < n;
    known += (size_t)cloud * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float *u = unknown + ((size_t)cloud * n + i) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += kTile) {
        const int len = min(kTile, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < len; t += kThreads) {
            const float *p = known + (size_t)(base + t) * 3;
            tile[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int t = 0; t < len; ++t) {
            const float4 p = tile[t];
            const float dx = ux - p.x;
            if (__fmul_rn(dx, dx) < b3) {
                const float d = pn2_sqdist(dx, uy - p.y, uz - p.z);
                const int k = base + t;
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else if (d < b3) {
                    b3 = d; i3 = k;
                }
            }
        }
    }
    if (active) {
        float *d = dist2 + ((size_t)cloud * n + i) * 3;
        int32_t *o = idx + ((size_t)cloud * n + i) * 3;
        d[0] = b1; d[1] = b2; d[2] = b3;
        o[0] = i1; o[1] = i2; o[2] = i3;
    }
}


// ---- culled three_nn (large levels) ----
// The unknown points are walked in Hilbert-curve order (spatial_order.cuh), so the 32 of a warp
// share a small bounding box.  Every warp works on its own: known points are streamed 256 at a
// time; before each pass the warp takes the largest current third-neighbour distance B of its
// lanes, and a known point can only improve some lane's top three if it lies inside the box grown
// by sqrt(B) (every axis term of the squared distance is non-negative and monotonically rounded;
// the box is grown by a further 0.1 % plus rounding slack).  The warp compacts the survivors with
// a ballot, in index order, and its lanes run the reference's strict-< insertion over the short
// list.  Ties are therefore resolved exactly as in the full scan and dist2 / idx are unchanged.
constexpr int kNnWarps = kThreads / 32;
constexpr int kNnList = 256;                     // known points a warp culls per pass (8 rounds of 32)

__global__ void __launch_bounds__(kThreads) three_nn_culled_kernel(const float *__restrict__ unknown,
                                                                  const float *__restrict__ known,
                                                                  const int32_t *__restrict__ order,
                                                                  float *__restrict__ dist2, int32_t *__restrict__ idx,
                                                                  int n, int m, int cpw) {
    __shared__ float4 cand[kNnWarps][kNnList];
    const int cloud = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // cpw unknown points per warp (lanes >= cpw are inactive but take part in the cull): see ball_query_culled_kernel
    const int wfirst = (blockIdx.x * kNnWarps + warp) * cpw;
    if (wfirst >= n) return;                     // warps never meet at a CTA barrier
    const bool active = lane < cpw && wfirst + lane < n;
    // inactive lanes mirror the warp's first point: they neither widen the box nor raise the bound
    const int i = __ldg(order + (size_t)cloud * n + (active ? wfirst + lane : wfirst));
    known += (size_t)cloud * m * 3;
    const float *u = unknown + ((size_t)cloud * n + i) * 3;
    const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);

    float lx = ux, hx = ux, ly = uy, hy = uy, lz = uz, hz = uz;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
        lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    // a NaN coordinate anywhere in the warp disables the cull (every comparison below would be false)
    const bool box_ok = !__any_sync(0xffffffffu, !(ux == ux) || !(uy == uy) || !(uz == uz)) && (lx <= hx) && (ly <= hy) &&
                        (lz <= hz);
    // rounding slack of the box arithmetic itself, relative to the coordinate magnitude
    const float ex = 1e-4f + 1e-6f * fmaxf(fabsf(lx), fabsf(hx)), ey = 1e-4f + 1e-6f * fmaxf(fabsf(ly), fabsf(hy)),
                ez = 1e-4f + 1e-6f * fmaxf(fabsf(lz), fabsf(hz));

    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    float4 *mine = cand[warp];
    for (int base = 0; base < m; base += kNnList) {
        // ---- bound: largest third-neighbour distance in the warp ----
        float bm = b3;
#pragma unroll
        for (int o = 16; o; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        const bool keep_all = !box_ok || !(bm < 1.0e30f);
        const float g = keep_all ? 0.f : sqrtf(bm) * 1.001f;
        const float clx = lx - (g + ex), chx = hx + (g + ex), cly = ly - (g + ey), chy = hy + (g + ey),
                    clz = lz - (g + ez), chz = hz + (g + ez);
        // ---- cull the next 256 known points, 32 per round, compacted in index order ----
        int wn = 0;
#pragma unroll
        for (int rd = 0; rd < kNnList / 32; ++rd) {
            const int k = base + rd * 32 + lane;
            bool in = false;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (k < m) {
                px = __ldg(known + (size_t)k * 3); py = __ldg(known + (size_t)k * 3 + 1); pz = __ldg(known + (size_t)k * 3 + 2);
                in = keep_all || (px >= clx && px <= chx && py >= cly && py <= chy && pz >= clz && pz <= chz);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) mine[wn + __popc(bal & ((1u << lane) - 1u))] = make_float4(px, py, pz, __int_as_float(k));
            wn += __popc(bal);
        }
        __syncwarp();
        // ---- the reference's insertion over the list; four distances are evaluated together so
        //      their latencies overlap, the (rarely taken) insertions stay in index order ----
        auto insert = [&](float d, int k) {
            if (d < b3) {
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else {
                    b3 = d; i3 = k;
                }
            }
        };
        int t = 0;
        for (; t + 4 <= wn; t += 4) {
            float dd[4];
            int kk[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 p = mine[t + v];
                dd[v] = pn2_sqdist(ux - p.x, uy - p.y, uz - p.z);
                kk[v] = __float_as_int(p.w);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) insert(dd[v], kk[v]);
        }
        for (; t < wn; ++t) {
            const float4 p = mine[t];
            insert(pn2_sqdist(ux - p.x, uy - p.y, uz - p.z), __float_as_int(p.w));
        }
        __syncwarp();
    }
    if (active) {
        float *d = dist2 + ((size_t)cloud * n + i) * 3;
        int32_t *o = idx + ((size_t)cloud * n + i) * 3;
        d[0] = b1; d[1] = b2; d[2] = b3;
        o[0] = i1; o[1] = i2; o[2] = i3;
    }
}

constexpr int kChan = 8;

__global__ void __launch_bounds__(256) three_interpolate_kernel(const float *__restrict__ points,
                                                               const int32_t *__restrict__ idx,
                                                               const float *__restrict__ weight,
                                                               float *__restrict__ out, int c, int m, int n) {
    const int cloud = blockIdx.z;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c0 = blockIdx.y * kChan;
    const int32_t *id = idx + ((size_t)cloud * n + i) * 3;
    const float *w = weight + ((size_t)cloud * n + i) * 3;
    const int k0 = __ldg(id), k1 = __ldg(id + 1), k2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *src = points + ((size_t)cloud * c + c0) * m;
    float *dst = out + ((size_t)cloud * c + c0) * n + i;
#pragma unroll
    for (int j = 0; j < kChan; ++j) {
        if (c0 + j < c) {
            const float *row = src + (size_t)j * m;
            float t = __fmul_rn(w1, __ldg(row + k1));
            t = __fmaf_rn(w0, __ldg(row + k0), t);
            t = __fmaf_rn(w2, __ldg(row + k2), t);
            dst[(size_t)j * n] = t;
        }
    }
}

__global__ void __launch_bounds__(256) three_interpolate_grad_kernel(const float *__restrict__ grad_out,
                                                                    const int32_t *__restrict__ idx,
                                                                    const float *__restrict__ weight,
                                                                    float *__restrict__ grad_points, int c, int n,
                                                                    int m) {
    const int cloud = blockIdx.z;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c0 = blockIdx.y * kChan;
    const int32_t *id = idx + ((size_t)cloud * n + i) * 3;
    const float *w = weight + ((size_t)cloud * n + i) * 3;
    const int k0 = __ldg(id), k1 = __ldg(id + 1), k2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *src = grad_out + ((size_t)cloud * c + c0) * n + i;
    float *dst = grad_points + ((size_t)cloud * c + c0) * m;
#pragma unroll
    for (int j = 0; j < kChan; ++j) {
        if (c0 + j < c) {
            const float g = __ldg(src + (size_t)j * n);
            float *row = dst + (size_t)j * m;
            atomicAdd(row + k0, g * w0);
            atomicAdd(row + k1, g * w1);
            atomicAdd(row + k2, g * w2);
        }
    }
}

}  // namespace

PN2_API int pn2_three_nn_f32(const float *unknown, const float *known, float *dist2, int32_t *idx, int b, int n, int m,
                             cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0) {
        pn2_set_last_error("pn2_three_nn_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || n == 0) return PN2_OK;
    dim3 grid(pn2_divup(n, kThreads), b);
    three_nn_kernel<<<grid, kThreads, 0, stream>>>(unknown, known, dist2, idx, n, m);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// pn2_three_nn_f32 through the spatially culled scan; `order` is caller scratch of b * n int32
// (left holding the Hilbert order of the unknown points).  order == NULL or a small level runs the
// brute-force kernel.  Same dist2 / idx bit for bit.
PN2_API int pn2_three_nn_culled_f32(const float *unknown, const float *known, float *dist2, int32_t *idx, int32_t *order,
                                    int b, int n, int m, cudaStream_t stream) {
    if (!order || n < 1024 || m < 512) return pn2_three_nn_f32(unknown, known, dist2, idx, b, n, m, stream);
    if (b < 0) {
        pn2_set_last_error("pn2_three_nn_culled_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    if (launch_spatial_order(unknown, order, b, n, stream) != cudaSuccess) {
        pn2_set_last_error("pn2_three_nn_culled_f32: ordering kernel launch failed");
        return PN2_ERR_LAUNCH;
    }
    static int cpw_env = -1;                      // PN2_NN_CPW: unknown points per warp (tuning; 32, 16 or 8)
    if (cpw_env < 0) {
        const char *e = getenv("PN2_NN_CPW");
        cpw_env = e ? atoi(e) : 0;
        if (cpw_env != 8 && cpw_env != 16 && cpw_env != 32) cpw_env = 0;
    }
    const int cpw = cpw_env ? cpw_env : 32;
    dim3 grid(pn2_divup(n, kNnWarps * cpw), b);
    three_nn_culled_kernel<<<grid, kThreads, 0, stream>>>(unknown, known, order, dist2, idx, n, m, cpw);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_three_interpolate_f32(const float *points, const int32_t *idx, const float *weight, float *out, int b,
                                      int c, int m, int n, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0) {
        pn2_set_last_error("pn2_three_interpolate_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || c == 0 || n == 0) return PN2_OK;
    dim3 grid(pn2_divup(n, 256), pn2_divup(c, kChan), b);
    three_interpolate_kernel<<<grid, 256, 0, stream>>>(points, idx, weight, out, c, m, n);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx, const float *weight,
                                           float *grad_points, int b, int c, int n, int m, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0) {
        pn2_set_last_error("pn2_three_interpolate_grad_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || c == 0 || n == 0) return PN2_OK;
    dim3 grid(pn2_divup(n, 256), pn2_divup(c, kChan), b);
    three_interpolate_grad_kernel<<<grid, 256, 0, stream>>>(grad_out, idx, weight, grad_points, c, n, m);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// ---- point-major 3-tap interpolation (internal layout of the fused FP path) ----
// feats (B, m, ldf) rows of C channels ; idx / weight (B, n, 3) ; out rows (B*n, ldo) cols [0,C).
// Same arithmetic order as the channel-major kernel; a row gather is three contiguous reads.
namespace {
__global__ void __launch_bounds__(256) three_interpolate_pm_kernel(const float *__restrict__ feats, int ldf,
                                                                  const int32_t *__restrict__ idx,
                                                                  const float *__restrict__ weight,
                                                                  float *__restrict__ out, int ldo, int c, int m,
                                                                  int n, long long total) {
    // one warp per unknown point, lanes stride over channels
    const long long pt = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= total) return;
    const long long cloud = pt / n;
    const int32_t *id = idx + pt * 3;
    const float *w = weight + pt * 3;
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *r0 = feats + (cloud * m + __ldg(id)) * ldf;
    const float *r1 = feats + (cloud * m + __ldg(id + 1)) * ldf;
    const float *r2 = feats + (cloud * m + __ldg(id + 2)) * ldf;
    float *o = out + pt * ldo;
    for (int ch = lane; ch < c; ch += 32) {
        float t = __fmul_rn(w1, __ldg(r1 + ch));
        t = __fmaf_rn(w0, __ldg(r0 + ch), t);
        t = __fmaf_rn(w2, __ldg(r2 + ch), t);
        o[ch] = t;
    }
}

// The same interpolation with the weights formed in the kernel from the SQUARED distances of three_nn: the statements
// the FP module runs in torch between the two ops (pointnet2_utils.py:104 sqrt, pointnet2_modules.py:209-211
// 1 / (dist + 1e-8), sum over the three, division) with the same IEEE operations in the same order.
template <bool VEC>
__global__ void __launch_bounds__(256) three_interpolate_pm_d2_kernel(const float *__restrict__ feats, int ldf,
                                                                     const int32_t *__restrict__ idx,
                                                                     const float *__restrict__ dist2,
                                                                     float *__restrict__ out, int ldo, int c, int m,
                                                                     int n, long long total, int sum_order) {
    const long long pt = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= total) return;
    const long long cloud = pt / n;
    // lanes 0..2 each form one reciprocal and one weight (a square root and two IEEE divisions per LANE instead of three and six
    // per warp-wide instruction stream: the first version spent two thirds of its instructions on these uniform values)
    const int l3 = lane < 3 ? lane : 0;
    const float q = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + pt * 3 + l3)), 1e-8f));
    const int id = __ldg(idx + pt * 3 + l3);
    const float q0 = __shfl_sync(0xffffffffu, q, 0), q1 = __shfl_sync(0xffffffffu, q, 1), q2 = __shfl_sync(0xffffffffu, q, 2);
    // association of the three-term sum: torch.sum's reduction kernel is not left-to-right (two threads take the even and
    // the odd elements), so the order is a parameter pinned against torch by the test
    const float norm = sum_order == 0 ? __fadd_rn(__fadd_rn(q0, q1), q2)
                     : sum_order == 1 ? __fadd_rn(__fadd_rn(q0, q2), q1) : __fadd_rn(q0, __fadd_rn(q1, q2));
    const float w = __fdiv_rn(q, norm);
    const float w0 = __shfl_sync(0xffffffffu, w, 0), w1 = __shfl_sync(0xffffffffu, w, 1), w2 = __shfl_sync(0xffffffffu, w, 2);
    const float *r0 = feats + (cloud * m + __shfl_sync(0xffffffffu, id, 0)) * ldf;
    const float *r1 = feats + (cloud * m + __shfl_sync(0xffffffffu, id, 1)) * ldf;
    const float *r2 = feats + (cloud * m + __shfl_sync(0xffffffffu, id, 2)) * ldf;
    float *o = out + pt * ldo;
    if (VEC) {
        for (int ch = lane * 4; ch < c; ch += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(r0 + ch)), b4 = __ldg(reinterpret_cast<const float4 *>(r1 + ch)),
                         c4 = __ldg(reinterpret_cast<const float4 *>(r2 + ch));
            float4 t;
            t.x = __fmaf_rn(w2, c4.x, __fmaf_rn(w0, a.x, __fmul_rn(w1, b4.x)));
            t.y = __fmaf_rn(w2, c4.y, __fmaf_rn(w0, a.y, __fmul_rn(w1, b4.y)));
            t.z = __fmaf_rn(w2, c4.z, __fmaf_rn(w0, a.z, __fmul_rn(w1, b4.z)));
            t.w = __fmaf_rn(w2, c4.w, __fmaf_rn(w0, a.w, __fmul_rn(w1, b4.w)));
            *reinterpret_cast<float4 *>(o + ch) = t;
        }
    } else {
        for (int ch = lane; ch < c; ch += 32) {
            float t = __fmul_rn(w1, __ldg(r1 + ch));
            t = __fmaf_rn(w0, __ldg(r0 + ch), t);
            t = __fmaf_rn(w2, __ldg(r2 + ch), t);
            o[ch] = t;
        }
    }
}
}  // namespace

// pn2_three_interpolate_pm_f32 fed with the squared distances of pn2_three_nn_f32 instead of ready-made weights.
// sum_order: 0 = (r0 + r1) + r2, 1 = (r0 + r2) + r1, 2 = r0 + (r1 + r2) for the normalising sum of the three reciprocals.
PN2_API int pn2_three_interpolate_pm_d2_f32(const float *feats, int ldf, const int32_t *idx, const float *dist2,
                                            float *out, int ldo, int b, int c, int m, int n, int sum_order,
                                            cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0 || ldf < c || ldo < c || sum_order < 0 || sum_order > 2) {
        pn2_set_last_error("pn2_three_interpolate_pm_d2_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    const long long total = (long long)b * n;
    if (total == 0 || c == 0) return PN2_OK;
    const bool vec = !(c & 3) && !(ldf & 3) && !(ldo & 3) && !(reinterpret_cast<uintptr_t>(feats) & 15) &&
                     !(reinterpret_cast<uintptr_t>(out) & 15);
    if (vec)
        three_interpolate_pm_d2_kernel<true><<<pn2_divup(total * 32, 256), 256, 0, stream>>>(feats, ldf, idx, dist2, out, ldo, c,
                                                                                            m, n, total, sum_order);
    else
        three_interpolate_pm_d2_kernel<false><<<pn2_divup(total * 32, 256), 256, 0, stream>>>(feats, ldf, idx, dist2, out, ldo, c,
                                                                                             m, n, total, sum_order);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_three_interpolate_pm_f32(const float *feats, int ldf, const int32_t *idx, const float *weight,
                                         float *out, int ldo, int b, int c, int m, int n, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0 || ldf < c || ldo < c) {
        pn2_set_last_error("pn2_three_interpolate_pm_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    const long long total = (long long)b * n;
    if (total == 0 || c == 0) return PN2_OK;
    three_interpolate_pm_kernel<<<pn2_divup(total * 32, 256), 256, 0, stream>>>(feats, ldf, idx, weight, out, ldo, c, m,
                                                                               n, total);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
