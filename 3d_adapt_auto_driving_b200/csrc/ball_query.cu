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
#include "spatial_order.cuh"
#include <cstdlib>

namespace {

constexpr int kThreads = 128;   // centres per CTA
constexpr int kTile = 1024;     // candidates per shared-memory tile (16 KB)

template <bool DUAL>
__global__ void __launch_bounds__(kThreads) ball_query_kernel(const float *__restrict__ new_xyz,
                                                             const float *__restrict__ xyz, int32_t *__restrict__ idx0,
                                                             int32_t *__restrict__ idx1, int n, int m, float radius0,
                                                             int ns0, float radius1, int ns1, int32_t *__restrict__ hits0,
                                                             int32_t *__restrict__ hits1) {
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
    // optional: number of neighbours found (capped at nsample), for the duplicate-skipping SA kernels
    if (active && hits0) hits0[(size_t)cloud * m + c] = min(cnt0, ns0);
    if (DUAL && active && hits1) hits1[(size_t)cloud * m + c] = min(cnt1, ns1);
}


// =================================================================================================
// Culled variant (large clouds).  The brute-force scan above tests every (centre, point) pair; with
// KITTI extents (70 m x 80 m) and radii of 0.1 - 1 m more than 95 % of those tests fail on the
// first axis.  Here the centres of a cloud are first put in Hilbert-curve order of their ground-plane
// cell (spatial_order.cuh: one CTA per cloud, counting sort in shared memory), so that the 32
// centres of a WARP share a small bounding box.  Each warp works on its own: it streams the cloud
// 512 points at a time, compacts with a ballot the points inside its box grown by the larger
// radius (ascending index order is preserved), and its lanes then scan only that short list with
// the exact test of the reference.  Warps are the unit of work because lidar clouds are very
// unevenly dense: the few warps whose centres sit in the dense near field keep thousands of
// candidates, the rest a few hundred, and 128 small units per cloud balance over the SMs.
// The list is scanned cooperatively (scan_list): 32 candidates at a time against one centre, hits
// ordered by a ballot and appended with consecutive stores -- no divergent per-lane insertion,
// whose serial latency otherwise dominates in the dense warps.  The cull is conservative (box grown by r * 1.001 plus rounding slack; a hit needs
// |d| < r on every axis because the squared distance is a sum of non-negative, monotonically
// rounded terms), so the set and order of accepted candidates -- and hence idx -- is unchanged.
// =================================================================================================
constexpr int kCullThreads = 64;                  // 2 autonomous warps per CTA
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kWarpList = 1024;                   // capacity of a warp's candidate list (16 KB; 8 KB lists, twice the resident
                                                  // warps, measured slower: 0.66 vs 0.62 ms of ball query per step)
constexpr int kScanUnroll = 8;                    // groups of 32 list entries evaluated per scan step
constexpr int kCullBatch = 8;                     // rounds of 32 candidates whose loads are in flight together

// One pass of the cooperative scan: the 32 lanes test 32 list entries at a time against ONE centre,
// a ballot orders the hits, and they are appended to the centre's row with consecutive stores.
// cnt / first live in the lane that owns the centre and are exchanged by shuffle.
template <bool DUAL>
__device__ __forceinline__ void scan_list(const float4 *__restrict__ list, int wn, int lane, float qx, float qy, float qz,
                                          int c, float r2_0, float r2_1, int ns0, int ns1, int32_t *__restrict__ idx0,
                                          int32_t *__restrict__ idx1, size_t row_base, int &cnt0, int &cnt1, int &first0,
                                          int &first1) {
    const unsigned lt = (1u << lane) - 1u;
    for (int j = 0; j < 32; ++j) {
        int n0 = __shfl_sync(0xffffffffu, cnt0, j);
        int n1 = DUAL ? __shfl_sync(0xffffffffu, cnt1, j) : ns1;
        if (n0 >= ns0 && n1 >= ns1) continue;     // this centre is full (or an inactive lane)
        const float cx = __shfl_sync(0xffffffffu, qx, j), cy = __shfl_sync(0xffffffffu, qy, j),
                    cz = __shfl_sync(0xffffffffu, qz, j);
        const int cj = __shfl_sync(0xffffffffu, c, j);
        int32_t *row0 = idx0 + (row_base + cj) * ns0;
        int32_t *row1 = DUAL ? idx1 + (row_base + cj) * ns1 : nullptr;
        int f0 = __shfl_sync(0xffffffffu, first0, j), f1 = DUAL ? __shfl_sync(0xffffffffu, first1, j) : 0;
        // kScanUnroll groups of 32 candidates per step: their loads, distances and ballots are
        // independent and overlap; only the cheap integer appends are sequential
        for (int i0 = 0; i0 < wn; i0 += 32 * kScanUnroll) {
            unsigned b0[kScanUnroll], b1[kScanUnroll];
            int kk[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                const int i = i0 + u * 32 + lane;
                const bool live = i < wn;
                const float4 p = list[live ? i : 0];
                const float d2 = pn2_sqdist(cx - p.x, cy - p.y, cz - p.z);
                kk[u] = __float_as_int(p.w);
                b0[u] = __ballot_sync(0xffffffffu, live && d2 < r2_0);
                b1[u] = DUAL ? __ballot_sync(0xffffffffu, live && d2 < r2_1) : 0u;
            }
            // appends: the slot of a hit is n + (hits of earlier groups) + (hits of lower lanes), so
            // the stores are predicated and independent of each other -- no serial chain through n
            if (n0 < ns0) {
                if (n0 == 0) {               // first hit of this centre so far: remember it for the fill rule
                    unsigned fb = 0u;
                    int fk = 0;
#pragma unroll
                    for (int u = kScanUnroll - 1; u >= 0; --u)
                        if (b0[u]) { fb = b0[u]; fk = kk[u]; }
                    if (fb) f0 = __shfl_sync(0xffffffffu, fk, __ffs(fb) - 1);
                }
                int at = n0;
#pragma unroll
                for (int u = 0; u < kScanUnroll; ++u) {
                    const int pos = at + __popc(b0[u] & lt);
                    if (((b0[u] >> lane) & 1u) && pos < ns0) row0[pos] = kk[u];
                    at += __popc(b0[u]);
                }
                n0 = at;
            }
            if (DUAL && n1 < ns1) {
                if (n1 == 0) {
                    unsigned fb = 0u;
                    int fk = 0;
#pragma unroll
                    for (int u = kScanUnroll - 1; u >= 0; --u)
                        if (b1[u]) { fb = b1[u]; fk = kk[u]; }
                    if (fb) f1 = __shfl_sync(0xffffffffu, fk, __ffs(fb) - 1);
                }
                int at = n1;
#pragma unroll
                for (int u = 0; u < kScanUnroll; ++u) {
                    const int pos = at + __popc(b1[u] & lt);
                    if (((b1[u] >> lane) & 1u) && pos < ns1) row1[pos] = kk[u];
                    at += __popc(b1[u]);
                }
                n1 = at;
            }
            if (n0 >= ns0 && n1 >= ns1) break;
        }
        if (lane == j) {
            cnt0 = min(n0, ns0); first0 = f0;
            if (DUAL) { cnt1 = min(n1, ns1); first1 = f1; }
        }
    }
}

template <bool DUAL>
__global__ void __launch_bounds__(kCullThreads) ball_query_culled_kernel(const float *__restrict__ new_xyz,
                                                                        const float *__restrict__ xyz,
                                                                        const int32_t *__restrict__ order,
                                                                        int32_t *__restrict__ idx0,
                                                                        int32_t *__restrict__ idx1, int n, int m,
                                                                        float radius0, int ns0, float radius1, int ns1,
                                                                        int zero_empty, int32_t *__restrict__ hits0,
                                                                        int32_t *__restrict__ hits1, int cpw) {
    __shared__ float4 cand[kCullWarps][kWarpList];
    const int cloud = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // cpw centres per warp (lanes >= cpw stay inactive but help with the cull and the scan): fewer centres mean a smaller
    // box, shorter candidate lists and more warps to spread the sparse regions' long scans over
    const int wfirst = (blockIdx.x * kCullWarps + warp) * cpw;
    if (wfirst >= m) return;                      // warps never meet at a CTA barrier
    const bool active = lane < cpw && wfirst + lane < m;
    // inactive lanes mirror the warp's first centre so they do not widen the box
    const int slot = active ? wfirst + lane : wfirst;
    const int c = order ? __ldg(order + (size_t)cloud * m + slot) : slot;
    xyz += (size_t)cloud * n * 3;
    const float *q = new_xyz + ((size_t)cloud * m + c) * 3;
    const float qx = __ldg(q + 0), qy = __ldg(q + 1), qz = __ldg(q + 2);

    const float r2_0 = __fmul_rn(radius0, radius0);
    const float r2_1 = DUAL ? __fmul_rn(radius1, radius1) : 0.f;
    const float rmax = DUAL ? fmaxf(fabsf(radius0), fabsf(radius1)) : fabsf(radius0);
    const float grow = rmax * 1.001f;

    float lx = qx, hx = qx, ly = qy, hy = qy, lz = qz, hz = qz;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
        lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    {   // grown box, plus the rounding slack of the box arithmetic itself (relative to the coordinate magnitude)
        const float ex = grow + 1e-4f + 1e-6f * fmaxf(fabsf(lx), fabsf(hx)), ey = grow + 1e-4f + 1e-6f * fmaxf(fabsf(ly), fabsf(hy)),
                    ez = grow + 1e-4f + 1e-6f * fmaxf(fabsf(lz), fabsf(hz));
        lx -= ex; hx += ex; ly -= ey; hy += ey; lz -= ez; hz += ez;
    }
    // a NaN centre makes the box NaN and every cull test false: fall back to "keep everything"
    const bool keep_all = __any_sync(0xffffffffu, !(qx == qx) || !(qy == qy) || !(qz == qz)) || !(lx <= hx) ||
                          !(ly <= hy) || !(lz <= hz);

    const size_t row_base = (size_t)cloud * m;
    int cnt0 = active ? 0 : ns0, cnt1 = (DUAL && active) ? 0 : ns1, first0 = 0, first1 = 0;

    float4 *mine = cand[warp];
    int wn = 0;
    for (int base = 0; base < n; base += kCullBatch * 32) {
        // ---- cull: kCullBatch rounds of 32 candidates, all loads issued first; survivors are
        //      appended in index order ----
        float px[kCullBatch], py[kCullBatch], pz[kCullBatch];
#pragma unroll
        for (int rd = 0; rd < kCullBatch; ++rd) {
            const int k = min(base + rd * 32 + lane, n - 1);
            px[rd] = __ldg(xyz + (size_t)k * 3); py[rd] = __ldg(xyz + (size_t)k * 3 + 1); pz[rd] = __ldg(xyz + (size_t)k * 3 + 2);
        }
#pragma unroll
        for (int rd = 0; rd < kCullBatch; ++rd) {
            const int k = base + rd * 32 + lane;
            const bool in = k < n && (keep_all || (px[rd] >= lx && px[rd] <= hx && py[rd] >= ly && py[rd] <= hy &&
                                                   pz[rd] >= lz && pz[rd] <= hz));
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) mine[wn + __popc(bal & ((1u << lane) - 1u))] = make_float4(px[rd], py[rd], pz[rd], __int_as_float(k));
            wn += __popc(bal);
        }
        // ---- scan when the list cannot take another batch, or at the end of the cloud ----
        if (wn > kWarpList - kCullBatch * 32 || base + kCullBatch * 32 >= n) {
            __syncwarp();
            scan_list<DUAL>(mine, wn, lane, qx, qy, qz, c, r2_0, r2_1, ns0, ns1, idx0, idx1, row_base, cnt0, cnt1, first0,
                            first1);
            __syncwarp();
            wn = 0;
            const bool done = (cnt0 >= ns0) && (!DUAL || cnt1 >= ns1);
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
    if (active && hits0) hits0[row_base + c] = cnt0;          // neighbours found, capped at nsample (0: none)
    if (DUAL && active && hits1) hits1[row_base + c] = cnt1;
    // ---- fill rule of the reference: the first hit occupies every slot a later hit did not take ----
    for (int j = 0; j < 32; ++j) {
        if (!__shfl_sync(0xffffffffu, (int)active, j)) continue;
        const int cj = __shfl_sync(0xffffffffu, c, j);
        const int n0 = __shfl_sync(0xffffffffu, cnt0, j), f0 = __shfl_sync(0xffffffffu, first0, j);
        // (a centre without any hit keeps what the caller put there -- zeros, pointnet2_utils.py:177 -- unless the caller
        // asked for the zeros to be written here, which saves it the fill launch)
        if (n0 > 0 || zero_empty)
            for (int l = n0 + lane; l < ns0; l += 32) idx0[(row_base + cj) * ns0 + l] = n0 > 0 ? f0 : 0;
        if (DUAL) {
            const int n1 = __shfl_sync(0xffffffffu, cnt1, j), f1 = __shfl_sync(0xffffffffu, first1, j);
            if (n1 > 0 || zero_empty)
                for (int l = n1 + lane; l < ns1; l += 32) idx1[(row_base + cj) * ns1 + l] = n1 > 0 ? f1 : 0;
        }
    }
}

template <bool DUAL>
int launch_culled(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int32_t *order, int b, int n, int m,
                  float r0, int ns0, float r1, int ns1, int zero_empty, int32_t *hits0, int32_t *hits1, cudaStream_t stream) {
    // few centres per cloud: the ordering launch costs more than it saves, warps take the centres as they come
    const bool sorted = m >= 256;
    if (sorted && launch_spatial_order(new_xyz, order, b, m, stream) != cudaSuccess) {
        pn2_set_last_error("pn2_ball_query_culled_f32: ordering kernel launch failed");
        return PN2_ERR_LAUNCH;
    }
    if (!sorted) order = nullptr;
    static int cpw_env = -1;                      // PN2_BQ_CPW: centres per warp (tuning; 32, 16 or 8)
    if (cpw_env < 0) {
        const char *e = getenv("PN2_BQ_CPW");
        cpw_env = e ? atoi(e) : 0;
        if (cpw_env != 8 && cpw_env != 16 && cpw_env != 32) cpw_env = 0;
    }
    // measured on the bench clouds (B = 16, profiles/r2l_bench_cpw*.json): 32 centres per warp 0.88 ms of ball query per step, 16:
    // 0.63, 8: 0.70 (and 3.5 % fewer scenes/s with batches in flight: twice the cull passes); same lists for every value
    const int cpw = cpw_env ? cpw_env : 16;
    dim3 grid(pn2_divup(m, kCullWarps * cpw), b);   // 2 warps x cpw centres per CTA
    ball_query_culled_kernel<DUAL><<<grid, kCullThreads, 0, stream>>>(new_xyz, xyz, order, idx0, idx1, n, m, r0, ns0, r1, ns1, zero_empty,
                                                                      hits0, hits1, cpw);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

inline bool use_culled(const int32_t *order, int n, int m) { return order && n >= 128; }

}  // namespace

PN2_API int pn2_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n, int m, float radius,
                               int nsample, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0) {
        pn2_set_last_error("pn2_ball_query_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return PN2_OK;
    dim3 grid(pn2_divup(m, kThreads), b);
    ball_query_kernel<false><<<grid, kThreads, 0, stream>>>(new_xyz, xyz, idx, nullptr, n, m, radius, nsample, 0.f, 0, nullptr, nullptr);
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
                                                           nsample1, nullptr, nullptr);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// The same results as pn2_ball_query_f32 (nsample1 == 0) / pn2_ball_query_dual_f32 through the
// spatially culled scan.  `order` is caller-provided scratch of b * m int32 (the Hilbert order of
// the centres is left there when m >= 256); with order == NULL or a tiny cloud the brute-force kernels run.
static int ball_query_culled(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int32_t *order, int b, int n,
                             int m, float radius0, int nsample0, float radius1, int nsample1, int zero_empty, int32_t *hits0,
                             int32_t *hits1, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample0 <= 0 || nsample1 < 0 || (nsample1 > 0 && !idx1)) {
        pn2_set_last_error("pn2_ball_query_culled_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0) return PN2_OK;
    if (n == 0 || !use_culled(order, n, m)) {
        if (zero_empty) {      // the brute-force kernels only write hits: zero the lists on the stream (memset nodes, not launches)
            if (cudaMemsetAsync(idx0, 0, (size_t)b * m * nsample0 * sizeof(int32_t), stream) != cudaSuccess ||
                (nsample1 > 0 && cudaMemsetAsync(idx1, 0, (size_t)b * m * nsample1 * sizeof(int32_t), stream) != cudaSuccess)) {
                pn2_set_last_error("pn2_ball_query_culled_fill_f32: cudaMemsetAsync failed");
                return PN2_ERR_LAUNCH;
            }
        }
        if (n == 0) {
            if ((hits0 && cudaMemsetAsync(hits0, 0, (size_t)b * m * sizeof(int32_t), stream) != cudaSuccess) ||
                (hits1 && cudaMemsetAsync(hits1, 0, (size_t)b * m * sizeof(int32_t), stream) != cudaSuccess)) {
                pn2_set_last_error("pn2_ball_query_culled_fill_f32: cudaMemsetAsync failed");
                return PN2_ERR_LAUNCH;
            }
            return PN2_OK;
        }
        dim3 grid(pn2_divup(m, kThreads), b);
        if (nsample1 > 0)
            ball_query_kernel<true><<<grid, kThreads, 0, stream>>>(new_xyz, xyz, idx0, idx1, n, m, radius0, nsample0, radius1,
                                                                   nsample1, hits0, hits1);
        else
            ball_query_kernel<false><<<grid, kThreads, 0, stream>>>(new_xyz, xyz, idx0, nullptr, n, m, radius0, nsample0, 0.f, 0,
                                                                    hits0, nullptr);
        PN2_CHECK_LAUNCH();
        return PN2_OK;
    }
    if (nsample1 > 0)
        return launch_culled<true>(new_xyz, xyz, idx0, idx1, order, b, n, m, radius0, nsample0, radius1, nsample1, zero_empty,
                                   hits0, hits1, stream);
    return launch_culled<false>(new_xyz, xyz, idx0, nullptr, order, b, n, m, radius0, nsample0, 0.f, 0, zero_empty, hits0, nullptr,
                                stream);
}

PN2_API int pn2_ball_query_culled_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1,
                                      int32_t *order, int b, int n, int m, float radius0, int nsample0, float radius1,
                                      int nsample1, cudaStream_t stream) {
    return ball_query_culled(new_xyz, xyz, idx0, idx1, order, b, n, m, radius0, nsample0, radius1, nsample1, 0, nullptr, nullptr, stream);
}

// pn2_ball_query_culled_f32 for index lists the caller has NOT zeroed: the lists of centres without any neighbour are
// written as zeros here (the value the reference's pre-zeroed tensor keeps, pointnet2_utils.py:177), every other list
// is complete anyway.  Same results as zero-fill + pn2_ball_query_culled_f32, one launch less per list.
// hits0 / hits1 (B, M) int32 or NULL: the number of neighbours each centre found, capped at nsample (0 = none) -- what the
// duplicate-skipping SA kernels otherwise re-derive from the lists with a launch of their own (csrc/group_compact.cu).
PN2_API int pn2_ball_query_culled_fill_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1,
                                           int32_t *order, int b, int n, int m, float radius0, int nsample0, float radius1,
                                           int nsample1, int32_t *hits0, int32_t *hits1, cudaStream_t stream) {
    return ball_query_culled(new_xyz, xyz, idx0, idx1, order, b, n, m, radius0, nsample0, radius1, nsample1, 1, hits0, hits1,
                             stream);
}
