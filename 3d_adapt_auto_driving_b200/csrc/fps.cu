// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces pointrcnn/pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-253
// (furthest_point_sampling_kernel + launcher) behind pn2_fps_f32 (include/pn2_b200.h).
//
// Design (see DESIGN.md "FPS"): the reference runs one 1024-thread CTA per cloud and re-reads
// xyz + the running min-distance from global memory every round (20 B/point/round), with a
// 2*log2(bs)-barrier shared-memory tree per round.  Here a cloud is owned by a thread-block
// CLUSTER of up to 8 CTAs (8 x 256 threads).  Every thread keeps its P points and their running
// min-distance in registers for the whole kernel, so after the initial load the kernel touches
// HBM only to write one index per round.  Per round:
//   1. P distance updates per thread (FMA order identical to the reference),
//   2. warp argmax with two redux.sync (max on the float bits, min on the tie-break rank),
//   3. the winning lane of every warp publishes a 32-byte record {d2, rank, k, x, y, z} to ALL
//      CTAs of the cluster with st.async (DSMEM, completes on the destination's mbarrier),
//   4. every warp waits on its own CTA's mbarrier, reduces the CLUSTER*8 records redundantly
//      (again two redux.sync) and thereby knows the next centre without a second exchange.
// There is no __syncthreads and no cluster barrier inside the round loop.
//
// Round cost is issue slots + one exchange latency, so the per-point work is pared down:
//   * the P points of a thread sit in registers SORTED by the reference's tie-break rank (a one-time
//     sorting network at load), so "first strict maximum in register order" IS the reference's
//     in-thread winner and the update loop only carries (max, index): no second selection pass;
//   * the distance arithmetic runs two points at a time on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2,
//     each half IEEE round-to-nearest: bit-identical to the scalar sequence);
//   * the winner's coordinates come from a shared-memory copy of the CTA's points (one LDS.128 in one
//     lane) instead of being carried through the selection;
//   * the rank reduction (second redux) only runs when two lanes / records tie on the distance;
//   * the mbarrier wait is CTA-scoped: the records arrive through the async proxy (st.async +
//     complete_tx) straight into shared memory, and a cluster-scoped acquire made ptxas invalidate
//     the L1 (CCTL.IVALL) in every warp every round.
//
// Bit-exactness: the reference's winner among equal distances is decided by its launch shape:
// thread tid scans k = tid, tid+bs, ... keeping the FIRST maximum (strict >, sampling_gpu.cu:
// 136-137) and the tree (__update, :86-91) keeps the lower slot on ties at every level, which
// orders threads by the BIT-REVERSED thread id.  So the reference picks
//   max d2, then min bitrev_{log2 bs}(k mod bs), then min k      with bs = opt_n_threads(N).
// rank(k) = bitrev(k mod bs) * ceil(N/bs) + k / bs encodes the last two keys in one integer.
#include "common.cuh"
#include <cmath>

namespace {

constexpr int kThreads = 256;

struct __align__(16) FpsRecord {
    uint32_t d2bits;  // running min distance of the candidate (non-negative float: bits are monotone)
    uint32_t rinv;    // ~rank: larger is better, 0 = "no valid point"
    int32_t k;        // point index
    uint32_t pad;
    float x, y, z, w;  // its coordinates, so nobody has to fetch them
};
static_assert(sizeof(FpsRecord) == 32, "record is two 16-byte stores");

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pn2_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pn2_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(pn2_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     remote_addr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// tie-break rank of point k (see the header): bitrev(k mod bs) * ceil(N / bs) + k / bs
__device__ __forceinline__ uint32_t fps_rank(int k, int log2bs, int cnt) {
    const uint32_t tref = (uint32_t)k & ((1u << log2bs) - 1u);
    const uint32_t rev = log2bs ? (__brev(tref) >> (32 - log2bs)) : 0u;
    return rev * (uint32_t)cnt + ((uint32_t)k >> log2bs);
}

// P points per thread, CLUSTER CTAs per cloud.  grid = (CLUSTER, B).
// dynamic shared memory: P * kThreads float4 = this CTA's points, entry p * kThreads + tid = point kbase + p * kThreads
// THREADS: 256, or 64 for launches of many small clouds (see fps_launch)
template <int P, int CLUSTER, int THREADS = 256>
__global__ void __launch_bounds__(THREADS) fps_kernel(const float *__restrict__ xyz, float *__restrict__ temp,
                                                       int32_t *__restrict__ idx, int n, int m, int log2bs, int cnt,
                                                       const int32_t *__restrict__ viol, float *__restrict__ new_xyz) {
    constexpr int kThreads = THREADS, kWarps = THREADS / 32;   // (shadow the file-level defaults)
    constexpr int S = CLUSTER * kWarps;  // records per round
    constexpr int P2 = (P + 1) / 2;
    extern __shared__ float4 pts_s[];
    __shared__ FpsRecord slots[2][S];
    __shared__ __align__(8) uint64_t bars[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t crank = (CLUSTER > 1) ? cluster_ctarank() : 0u;
    const int cloud = blockIdx.y;

    xyz += (size_t)cloud * n * 3;
    idx += (size_t)cloud * m;
    if (temp) temp += (size_t)cloud * n;
    if (new_xyz) new_xyz += (size_t)cloud * m * 3;     // optional: coordinates of the picked points, (B, M, 3)
    // guarded launch (pn2_fps_guarded_f32): pn2_fps_prefix_check_f32 proved that this cloud's answer is 0, 1, ..., m-1.
    // The flag is per cloud, so every CTA of the cluster leaves here, before any cluster-scope operation.
    if (viol != nullptr && __ldg(viol + cloud) == 0) {
        for (int i = (int)crank * kThreads + tid; i < m; i += CLUSTER * kThreads) idx[i] = i;
        if (new_xyz)
            for (int i = (int)crank * kThreads + tid; i < 3 * m; i += CLUSTER * kThreads) new_xyz[i] = __ldg(xyz + i);
        return;
    }

    // ---- resident state: P points in tie-break-rank order, their running min distance and index ----
    const int cta_base = (int)crank * (P * kThreads);
    const int kbase = cta_base + tid;
    int ks[P];
    {
        uint32_t pr[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = kbase + p * kThreads;
            ks[p] = k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < n) {
                v.x = __ldg(xyz + (size_t)k * 3 + 0);
                v.y = __ldg(xyz + (size_t)k * 3 + 1);
                v.z = __ldg(xyz + (size_t)k * 3 + 2);
                pr[p] = fps_rank(k, log2bs, cnt);
            } else {
                pr[p] = 0xFFFFFFFFu;   // padding: worst rank, sorts last
            }
            pts_s[p * kThreads + tid] = v;
        }
        // odd-even transposition network on (rank, k): fully unrolled, static register indices
#pragma unroll
        for (int pass = 0; pass < P; ++pass) {
#pragma unroll
            for (int i = pass & 1; i + 1 < P; i += 2) {
                const bool sw = pr[i] > pr[i + 1];
                const uint32_t a = pr[i], b = pr[i + 1];
                const int ka = ks[i], kb = ks[i + 1];
                pr[i] = sw ? b : a; pr[i + 1] = sw ? a : b;
                ks[i] = sw ? kb : ka; ks[i + 1] = sw ? ka : kb;
            }
        }
    }
    __syncthreads();   // pts_s complete (read back below in a different order, later by the winner lanes)
    float2 px[P2], py[P2], pz[P2], pt[P2];
#pragma unroll
    for (int s = 0; s < P; ++s) {
        const int k = ks[s];
        const float4 v = pts_s[k - cta_base];
        float t = 0.f;   // padding slot: distance 0 and the worst rank, can never win against a real point
        if (k < n) t = temp ? temp[k] : 1e10f;
        if (s & 1) { px[s >> 1].y = v.x; py[s >> 1].y = v.y; pz[s >> 1].y = v.z; pt[s >> 1].y = t; }
        else       { px[s >> 1].x = v.x; py[s >> 1].x = v.y; pz[s >> 1].x = v.z; pt[s >> 1].x = t; }
    }
    if (P & 1) { px[P2 - 1].y = 0.f; py[P2 - 1].y = 0.f; pz[P2 - 1].y = 0.f; pt[P2 - 1].y = 0.f; }

    if (CLUSTER > 1) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();  // barriers initialised everywhere before any remote store
    }

    float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);  // idx[0] = 0 always
    if (crank == 0 && tid == 0) {
        idx[0] = 0;
        if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
    }

    // cluster addresses of this warp's record slot and of the round barrier in every CTA (parity 0)
    uint32_t rslot[CLUSTER], rbar[CLUSTER];
#pragma unroll
    for (int cta = 0; cta < CLUSTER; ++cta) {
        rslot[cta] = CLUSTER > 1 ? mapa(pn2_smem_u32(&slots[0][crank * kWarps + warp]), cta) : 0u;
        rbar[cta] = CLUSTER > 1 ? mapa(pn2_smem_u32(&bars[0]), cta) : 0u;
    }

    for (int r = 0; r < m - 1; ++r) {
        const int par = r & 1;
        if (CLUSTER > 1 && tid == 0) mbar_arrive_expect_tx(&bars[par], S * (uint32_t)sizeof(FpsRecord));

        // 1. distance update and thread argmax: the first strict maximum in rank order
        float tmax = -1.f;
        int bestk = ks[0];
        const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
        for (int q = 0; q < P2; ++q) {
            // (p - c) == p + (-c) exactly; t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t) as pn2_sqdist
            const float2 dx = __fadd2_rn(px[q], ncx), dy = __fadd2_rn(py[q], ncy), dz = __fadd2_rn(pz[q], ncz);
            float2 t = __fmul2_rn(dy, dy);
            t = __ffma2_rn(dx, dx, t);
            t = __ffma2_rn(dz, dz, t);
            pt[q].x = fminf(t.x, pt[q].x);
            if (pt[q].x > tmax) { tmax = pt[q].x; bestk = ks[2 * q]; }
            if (2 * q + 1 < P) {
                pt[q].y = fminf(t.y, pt[q].y);
                if (pt[q].y > tmax) { tmax = pt[q].y; bestk = ks[2 * q + 1]; }
            }
        }
        // 2. warp argmax: max distance bits; the rank decides only if several lanes hold it
        const uint32_t vb = __float_as_uint(tmax);
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, vb);
        const uint32_t rk = (vb == wmax && bestk < n) ? fps_rank(bestk, log2bs, cnt) : 0xFFFFFFFFu;
        const uint32_t cand = __ballot_sync(0xffffffffu, vb == wmax);
        uint32_t winners = cand;
        if (cand & (cand - 1u)) {
            const uint32_t wrk = __reduce_min_sync(0xffffffffu, rk);
            winners = __ballot_sync(0xffffffffu, rk == wrk);
        }
        // 3. publish the warp's record
        if (lane == __ffs(winners) - 1) {
            const float4 c = pts_s[bestk - cta_base];
            const uint32_t rinv = ~rk;
            if (CLUSTER == 1) {
                FpsRecord rec;
                rec.d2bits = wmax; rec.rinv = rinv; rec.k = bestk; rec.pad = 0;
                rec.x = c.x; rec.y = c.y; rec.z = c.z; rec.w = 0.f;
                slots[par][warp] = rec;
            } else {
                const uint32_t so = (uint32_t)par * (uint32_t)(S * sizeof(FpsRecord)), bo = (uint32_t)par * 8u;
#pragma unroll
                for (int cta = 0; cta < CLUSTER; ++cta) {
                    const uint32_t dst = rslot[cta] + so;
                    const uint32_t dbar = rbar[cta] + bo;
                    st_async_v4(dst, wmax, rinv, (uint32_t)bestk, 0u, dbar);
                    st_async_v4(dst + 16, __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), 0u, dbar);
                }
            }
        }
        if (CLUSTER == 1) __syncthreads();
        else mbar_wait(&bars[par], (uint32_t)(r >> 1) & 1u);

        // 4. every warp reduces the S records (redundantly) -> next centre
        unsigned long long kA = 0ull, kB = 0ull;
        if (lane < S) kA = *reinterpret_cast<const unsigned long long *>(&slots[par][lane]);   // (rinv | d2bits << 32)
        if (S > 32 && lane + 32 < S) kB = *reinterpret_cast<const unsigned long long *>(&slots[par][lane + 32]);
        kA = (kA << 32) | (kA >> 32);
        kB = (kB << 32) | (kB >> 32);
        const unsigned long long km = kA > kB ? kA : kB;
        const int myslot = (kA >= kB) ? lane : lane + 32;
        const uint32_t hi = (uint32_t)(km >> 32), lo = (uint32_t)km;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
        uint32_t who = __ballot_sync(0xffffffffu, hi == mh);
        if (who & (who - 1u)) {
            const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            who = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
        }
        const int wslot = __shfl_sync(0xffffffffu, myslot, __ffs(who) - 1);
        const float4 c4 = *reinterpret_cast<const float4 *>(&slots[par][wslot].x);
        cx = c4.x; cy = c4.y; cz = c4.z;
        if (crank == 0 && tid == 0) {
            idx[r + 1] = slots[par][wslot].k;
            if (new_xyz) { new_xyz[3 * r + 3] = c4.x; new_xyz[3 * r + 4] = c4.y; new_xyz[3 * r + 5] = c4.z; }
        }
    }

    if (temp) {  // the reference leaves the running min distances in the caller's scratch
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int k = ks[s];
            if (k < n) temp[k] = (s & 1) ? pt[s >> 1].y : pt[s >> 1].x;
        }
    }
    if (CLUSTER > 1) cluster_sync_all();  // nobody exits while a peer may still target its smem
}

template <int P, int CLUSTER, int THREADS = 256>
cudaError_t launch_fps(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int log2bs, int cnt,
                       const int32_t *viol, float *new_xyz, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CLUSTER, b, 1);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)P * THREADS * sizeof(float4);
    if (cfg.dynamicSmemBytes > 48 * 1024) {
        static bool attr_done = false;   // per <P, CLUSTER> instantiation
        if (!attr_done) {
            cudaError_t ea = cudaFuncSetAttribute(fps_kernel<P, CLUSTER, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)cfg.dynamicSmemBytes);
            if (ea != cudaSuccess) return ea;
            attr_done = true;
        }
    }
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fps_kernel<P, CLUSTER, THREADS>, xyz, temp, idx, n, m, log2bs, cnt, viol, new_xyz);
}

template <int CLUSTER>
cudaError_t dispatch_p(int p, const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int log2bs, int cnt,
                       const int32_t *viol, float *new_xyz, cudaStream_t s) {
    switch (p) {
        case 1: return launch_fps<1, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
        case 2: return launch_fps<2, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
        case 4: return launch_fps<4, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
        case 8: return launch_fps<8, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
        case 16: return launch_fps<16, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
        default: return launch_fps<32, CLUSTER>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, s);
    }
}

// ---------------------------------------------------------------------------------------------
// "Is the answer simply 0, 1, ..., m-1?"  In the PointNet++ backbone every SA level samples the centres the level
// before it produced, i.e. a cloud that is ALREADY in FPS order, and FPS restricted to a prefix of an FPS ordering picks
// that prefix again: with the first r points picked, point r is the arg-max of the running min-distance over the whole
// parent cloud, hence over the subset.  The only way the 4095-round latency chain can give anything else is a TIE at the
// maximum, whose winner depends on the launch shape of the reference (see the header).  That is checked exactly, for
// any input and in parallel instead of in sequence:
//     D[k]   = min_{i<k} d2(p_k, p_i)                       (running min-distance of point k when it is due)
//     R_j(k) = min(1e10, min_{i<k} d2(p_j, p_i))  <  D[k]    for all k < m, j > k;   D[k] > 0 covers the picked j < k
// with d2 the reference's own expression (pn2_sqdist) -- the same floats the sequential kernel would compare.  Where the
// inequality is strict the arg-max of round k-1 is k whatever the tie order; an exact tie R_j(k) == D[k] is decided like the
// reference decides it, by the rank of the header of this file (k must have the lower rank).  Then idx = arange(m) is the
// reference's answer bit for bit; any violation (a lost tie, a larger candidate, duplicates of picked points, NaNs) bumps
// viol[cloud] and the guarded launch runs the real kernel for that cloud.  N * m pair evaluations, no dependence between them: ~20 us instead of 0.5 ms for 16 x (4096 -> 1024).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fps_prefix_dist_kernel(const float *__restrict__ xyz, float *__restrict__ dmin,
                                                              int32_t *__restrict__ viol, int n, int m) {
    // one warp per k in [1, m): D[k] = min over i < k
    const int cloud = blockIdx.y;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5) + 1;
    const int lane = threadIdx.x & 31;
    if (k >= m) return;
    const float *p = xyz + (size_t)cloud * n * 3;
    const float kx = __ldg(p + 3 * k), ky = __ldg(p + 3 * k + 1), kz = __ldg(p + 3 * k + 2);
    float t = 1e10f;
    for (int i = lane; i < k; i += 32) {
        const float cx = __ldg(p + 3 * i), cy = __ldg(p + 3 * i + 1), cz = __ldg(p + 3 * i + 2);
        t = fminf(pn2_sqdist(__fadd_rn(kx, -cx), __fadd_rn(ky, -cy), __fadd_rn(kz, -cz)), t);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) {
        dmin[(size_t)cloud * m + k] = t;
        if (!(t > 0.f)) atomicAdd(viol + cloud, 1);
    }
}

__global__ void __launch_bounds__(256) fps_prefix_check_kernel(const float *__restrict__ xyz, const float *__restrict__ dmin,
                                                               int32_t *__restrict__ viol, int n, int m, int log2bs, int cnt) {
    extern __shared__ float4 cen[];          // entry i < m-1: (p_i, D[i+1])
    const int cloud = blockIdx.y;
    const float *p = xyz + (size_t)cloud * n * 3;
    for (int i = threadIdx.x; i < m - 1; i += 256)
        cen[i] = make_float4(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2), __ldg(dmin + (size_t)cloud * m + i + 1));
    __syncthreads();
    const int j = blockIdx.x * 256 + threadIdx.x;
    bool bad = false;
    if (j < n && j >= 2) {
        const float jx = __ldg(p + 3 * j), jy = __ldg(p + 3 * j + 1), jz = __ldg(p + 3 * j + 2);
        const int last = min(j - 1, m - 1);          // rounds i = 0 .. last-1: the pick due after round i is i+1 < j
        float r = 1e10f;
        const uint32_t rank_j = fps_rank(j, log2bs, cnt);
        for (int i = 0; i < last; ++i) {
            const float4 c = cen[i];
            r = fminf(pn2_sqdist(__fadd_rn(jx, -c.x), __fadd_rn(jy, -c.y), __fadd_rn(jz, -c.z)), r);
            // an exact tie with the due point i + 1 is decided by the reference's rank (header of this file): only a tie
            // that j wins breaks the prefix
            bad |= !(r < c.w) && !(r == c.w && rank_j > fps_rank(i + 1, log2bs, cnt));
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(viol + cloud, 1);
}

}  // namespace

// fps_cells.cu: the pruned one-CTA kernel (clouds of up to 16384 points)
bool pn2_fps_cells_supported(int n);
cudaError_t pn2_fps_cells_launch(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int warps,
                                 const int32_t *viol, float *new_xyz, cudaStream_t stream);

// Block size the reference launcher would have used (cuda_utils.h:10-14); it only matters
// here because it fixes the tie-break order.  Same double-precision expression.
PN2_API int pn2_fps_ref_block_size(int n) {
    const int pow_2 = (int)(std::log((double)n) / std::log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

// xyz (B,N,3) f32 ; temp (B,N) f32 scratch or NULL (NULL = implicit 1e10, nothing written
// back) ; idx (B,M) int32 ; cluster_size: CTAs per cloud (0 = heuristic; 1, 2, 4, 8 force it: tests / tuning) ;
// viol: NULL, or (B) int32 from pn2_fps_prefix_check_f32 -- clouds with viol == 0 get idx = 0..M-1 without the
// round loop.  Enqueues on `stream`, never synchronises.
static int fps_launch(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int cluster_size,
                      const int32_t *viol, float *new_xyz, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || (!xyz && b * n > 0) || (!idx && b * m > 0)) {
        pn2_set_last_error("pn2_fps_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0) return PN2_OK;
    if (n == 0) {
        pn2_set_last_error("pn2_fps_f32: empty cloud with m > 0");
        return PN2_ERR_INVALID;
    }
    // Heuristic launches of mid-sized clouds go to the pruned one-CTA kernel (fps_cells.cu): a shorter round on a quarter
    // of the SMs.  Below kCellsMinN points the whole cloud is a handful of cells and the plain kernel's round is as short;
    // with few rounds the Hilbert-sort prepass (~20 us) does not pay.
    constexpr int kCellsMinN = 2048, kCellsMinM = 128;
    if (cluster_size == 0 && n > kCellsMinN && m >= kCellsMinM && pn2_fps_cells_supported(n)) {
        const cudaError_t ec = pn2_fps_cells_launch(xyz, temp, idx, b, n, m, 0, viol, new_xyz, stream);
        if (ec != cudaSuccess) {
            pn2_set_last_error(cudaGetErrorString(ec));
            return PN2_ERR_LAUNCH;
        }
        return PN2_OK;
    }
    const int bs = pn2_fps_ref_block_size(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    const int cnt = (n + bs - 1) / bs;

    // CTAs per cloud.  Measured (16384 -> 4096, v2 kernel): 4-CTA clusters 2.74 ms at B = 8 and B = 16;
    // 8-CTA clusters 3.52 ms at B = 8 (twice the records and remote stores per round for 8 instead of 16
    // points per thread) and 4.54 ms at B = 16 (an 8-CTA cluster must sit inside one GPC, 16 of them do
    // not fit in one wave); one CTA is best up to 4096 points (0.53 ms vs 0.62 ms for 4096 -> 1024).
    int cluster = n > 4096 ? 4 : 1;
    if (cluster_size) cluster = cluster_size;
    int p = (n + cluster * kThreads - 1) / (cluster * kThreads);
    while (p > 32 && cluster < 8) {
        cluster *= 2;
        p = (n + cluster * kThreads - 1) / (cluster * kThreads);
    }
    if (p > 32) {
        pn2_set_last_error("pn2_fps_f32: N > 65536 points per cloud is not supported");
        return PN2_ERR_UNSUPPORTED;
    }
    int pp = 1;
    while (pp < p) pp *= 2;
    cudaError_t e;
    // Many small clouds (the RCNN stage samples 1600 ROI clouds of 512 points): the launch is issue-bound, not latency-bound
    // (ncu: 74 % of the issue slots at 8 warps per cloud, most of them the per-warp argmax / record / reduce of each round),
    // so a cloud gets TWO warps with 4x the points per lane: ~3.5x fewer instructions per round.
    if (cluster == 1 && cluster_size == 0 && n <= 1024 && n > 64 && b >= 256) {
        int p64 = 1;
        while (p64 * 64 < n) p64 *= 2;
        switch (p64) {
            case 2: e = launch_fps<2, 1, 64>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
            case 4: e = launch_fps<4, 1, 64>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
            case 8: e = launch_fps<8, 1, 64>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
            default: e = launch_fps<16, 1, 64>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
        }
        if (e != cudaSuccess) {
            pn2_set_last_error(cudaGetErrorString(e));
            return PN2_ERR_LAUNCH;
        }
        return PN2_OK;
    }
    switch (cluster) {
        case 1: e = dispatch_p<1>(pp, xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
        case 2: e = dispatch_p<2>(pp, xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
        case 4: e = dispatch_p<4>(pp, xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
        case 8: e = dispatch_p<8>(pp, xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream); break;
        default: pn2_set_last_error("pn2_fps_f32: cluster must be 1, 2, 4 or 8"); return PN2_ERR_INVALID;
    }
    if (e != cudaSuccess) {
        pn2_set_last_error(cudaGetErrorString(e));
        return PN2_ERR_LAUNCH;
    }
    return PN2_OK;
}

PN2_API int pn2_fps_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, cudaStream_t stream) {
    return fps_launch(xyz, temp, idx, b, n, m, 0, nullptr, nullptr, stream);
}

// pn2_fps_f32 with the CTAs-per-cloud cluster size forced (1, 2, 4 or 8; 0 = heuristic).  Same result for every value.
PN2_API int pn2_fps_cluster_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int cluster_size,
                                cudaStream_t stream) {
    return fps_launch(xyz, temp, idx, b, n, m, cluster_size, nullptr, nullptr, stream);
}

// viol (B) int32 (zeroed here, on the stream); dmin (B, M) f32 scratch.  Afterwards viol[c] == 0 iff furthest point sampling of
// cloud c provably returns 0, 1, ..., M-1 (see fps_prefix_check_kernel); M <= 4096.
PN2_API int pn2_fps_prefix_check_f32(const float *xyz, float *dmin, int32_t *viol, int b, int n, int m, cudaStream_t stream) {
    if (b < 0 || n <= 0 || m <= 0 || m > n || !xyz || !dmin || !viol) {
        pn2_set_last_error("pn2_fps_prefix_check_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (m > 4096) {
        pn2_set_last_error("pn2_fps_prefix_check_f32: m > 4096 is not supported");
        return PN2_ERR_UNSUPPORTED;
    }
    if (b == 0) return PN2_OK;
    if (cudaMemsetAsync(viol, 0, (size_t)b * sizeof(int32_t), stream) != cudaSuccess) {   // a memset node, not a launch
        pn2_set_last_error("pn2_fps_prefix_check_f32: cudaMemsetAsync failed");
        return PN2_ERR_LAUNCH;
    }
    if (m == 1) return PN2_OK;
    fps_prefix_dist_kernel<<<dim3((m - 1 + 7) / 8, b), 256, 0, stream>>>(xyz, dmin, viol, n, m);
    PN2_CHECK_LAUNCH();
    const int bs = pn2_fps_ref_block_size(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    fps_prefix_check_kernel<<<dim3((n + 255) / 256, b), 256, (size_t)m * sizeof(float4), stream>>>(xyz, dmin, viol, n, m, log2bs,
                                                                                                  (n + bs - 1) / bs);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// pn2_fps_f32 (temp = NULL) guarded by pn2_fps_prefix_check_f32's verdict: clouds with viol == 0 skip the round loop.
PN2_API int pn2_fps_guarded_f32(const float *xyz, int32_t *idx, const int32_t *viol, int b, int n, int m,
                                cudaStream_t stream) {
    if (!viol) {
        pn2_set_last_error("pn2_fps_guarded_f32: viol is required");
        return PN2_ERR_INVALID;
    }
    return fps_launch(xyz, nullptr, idx, b, n, m, 0, viol, nullptr, stream);
}

// pn2_fps_f32 (temp = NULL) that also writes the coordinates of the picked points, new_xyz (B, M, 3): what the callers of
// furthest_point_sample do next with a gather (pointnet2_modules.py:29-33) costs the kernel three stores per round.
PN2_API int pn2_fps_xyz_f32(const float *xyz, int32_t *idx, float *new_xyz, int b, int n, int m, cudaStream_t stream) {
    if (!new_xyz && b * m > 0) {
        pn2_set_last_error("pn2_fps_xyz_f32: new_xyz is required");
        return PN2_ERR_INVALID;
    }
    return fps_launch(xyz, nullptr, idx, b, n, m, 0, nullptr, new_xyz, stream);
}

// pn2_fps_guarded_f32 + the coordinates of the picked points (clouds with viol == 0: the first M rows of xyz).
PN2_API int pn2_fps_guarded_xyz_f32(const float *xyz, int32_t *idx, float *new_xyz, const int32_t *viol, int b, int n, int m,
                                    cudaStream_t stream) {
    if (!viol || (!new_xyz && b * m > 0)) {
        pn2_set_last_error("pn2_fps_guarded_xyz_f32: viol and new_xyz are required");
        return PN2_ERR_INVALID;
    }
    return fps_launch(xyz, nullptr, idx, b, n, m, 0, viol, new_xyz, stream);
}
