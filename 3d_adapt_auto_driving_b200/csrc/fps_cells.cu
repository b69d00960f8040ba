// fps_cells.cu -- furthest point sampling with exact spatial pruning: one CTA per cloud, 2048 < N <= 16384.
//
// Replaces pointrcnn/pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-253 like fps.cu does (same result, bit for
// bit: max d2, then the reference's tree order among equal distances, see the header of fps.cu).
//
// Why a second kernel.  fps.cu updates the running min-distance of EVERY point in every round, which costs ~470 cycles of
// a ~1310-cycle round, and needs a 4-CTA cluster (with a ~840-cycle DSMEM exchange) to hold the cloud in registers.
// But a new centre c only lowers min-distances inside the ball around c whose radius is the current maximum: after a
// few hundred rounds that is a small neighbourhood.  Here the cloud is sorted along a Hilbert curve (in-kernel counting
// sort) and cut into CELLS of 128 consecutive points = 4 register slots x 32 lanes of one warp; a warp owns CPW cells
// that lie far apart on the curve.  Per cell the warp keeps the bounding box and the exact maximum `cmax` of the cell's
// running min-distances.  In a round a cell is touched only if
//      lb(c, box) < cmax ,    lb = sqdist(max(0, lo - c, c - hi))  evaluated with the reference's own float expression.
// Rounding is monotone, so lb is a true lower bound of the FLOAT distance the reference computes for every point of
// the cell:  d2(p, c) >= lb >= cmax >= t[p]  =>  min(t[p], d2) == t[p] -- skipping the cell changes nothing, the result is
// bit-identical.  On KITTI-shaped clouds ~3 of the 128 cells are touched per round (in-kernel stopwatch, tools/prof_fps_cells.py;
// the numpy model written before the kernel, tools/sim_fps_cells.py, said 3-5).
// Everything is on one SM, so the exchange is one __syncthreads per round:
//   1. every warp tests its CPW boxes (lane i < CPW tests cell i; one ballot, one redux for the first touched cell);
//   2. a touched cell: 3 x LDS.128 (coordinates live in shared memory, SoA, slot-major per lane; requested before the branch
//      tree that selects the cell's registers) + packed fp32 update of 4 running distances per lane, redux.max -> new cmax,
//      and the cell's RECORD (maximum distance bits, shared-memory position of the point that holds it) written to the
//      record buffer of this round; after the barrier the lane that owns the cell copies it into the other buffer too
//      (untouched cells therefore need no per-round work at all);
//   3. __syncthreads; every warp reduces the 128 cell records redundantly (four per lane, two redux.sync) and reads the next
//      centre's coordinates from shared memory.
// An exact tie on the distance -- inside a cell or between cells -- leaves these fast paths: the reference's rank decides, read
// from a u16 table of inverted ranks per shared-memory position (the point index, needed only for the output, is recovered
// from the rank after the loop).
// Registers hold only the running distances (4 * CPW per lane) -- 8 warps x 16 cells cover 16384 points in ONE CTA: 1.6 ms and
// 16 SMs for 16 x (16384 -> 4096) against 2.9 ms on 64 SMs for the 4-CTA cluster kernel (tools/bench_fps_cluster.py).
#include "spatial_order.cuh"
#include <cmath>
#include <type_traits>

namespace {

constexpr int kCellPts = 128;
constexpr int kMaxCells = 128;      // cells per cloud (16384 points); the record arrays always have this many entries

__device__ __forceinline__ uint32_t cells_rank(uint32_t k, int log2bs, int cnt) {
    const uint32_t tref = k & ((1u << log2bs) - 1u);
    const uint32_t rev = log2bs ? (__brev(tref) >> (32 - log2bs)) : 0u;
    return rev * (uint32_t)cnt + (k >> log2bs);
}

// explicit shared-memory accesses on 32-bit shared addresses (see the main loop)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

__device__ __forceinline__ void sts_u32_if(bool p, uint32_t a, uint32_t v) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.shared.u32 [%0], %1;\n}" ::"r"(a), "r"(v), "r"((uint32_t)p) : "memory");
}

// The u16 table of the main loop holds, per shared-memory position, the INVERTED reference rank of the point (~rank & 0xFFFF,
// larger is better; 0 = padding slot; real ranks are < N + 1024 <= 17408, so real entries are >= 48128).  A tie costs one
// table read per tied candidate; the point index, needed only for the output, is recovered from the rank.
__device__ __forceinline__ uint32_t rinv_of(uint32_t k, int log2bs, int cnt) { return ~cells_rank(k, log2bs, cnt) & 0xFFFFu; }
__device__ __forceinline__ uint32_t k_of_rinv(uint32_t rinv, int log2bs, int cnt) {
    const uint32_t rank = ~rinv & 0xFFFFu;
    if (log2bs == 0) return rank;
    const uint32_t rev = rank / (uint32_t)cnt, kd = rank - rev * (uint32_t)cnt;
    return (kd << log2bs) | (__brev(rev) >> (32 - log2bs));
}

// Exact paths (ties on the distance): called by the whole warp.  Out of line: inlined into the sixteen cell bodies they grew the
// round loop enough to slow EVERY phase of it by ~10 % (instruction fetch), ties or not.
// The slots of one cell that hold its maximum mm, over all candidate lanes: position of the one with the best rank.
__device__ __noinline__ uint32_t cell_exact_pos(bool cand, float p0, float p1, float p2, float p3, float mm, uint32_t posc,
                                                   uint32_t sK) {
    uint32_t best = 0u;
    if (cand) {
        const float p[4] = {p0, p1, p2, p3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (p[q] == mm) {
                const uint32_t key = (lds_u16(sK + (posc + q) * 2u) << 16) | (posc + q);
                best = key > best ? key : best;
            }
        }
    }
    return __reduce_max_sync(0xffffffffu, best) & 0xFFFFu;
}
// The cell records (four per lane) whose distance equals the maximum mh: position of the candidate with the best rank.
__device__ __noinline__ uint32_t records_exact_pos(uint4 h4, uint4 p4, uint32_t mh, uint32_t sK) {
    const uint32_t h[4] = {h4.x, h4.y, h4.z, h4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
    uint32_t best = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (h[j] == mh) {
            const uint32_t key = ((lds_u16(sK + p[j] * 2u) << 16) | p[j]) + 1u;
            best = key > best ? key : best;
        }
    }
    return (__reduce_max_sync(0xffffffffu, best) - 1u) & 0xFFFFu;
}

// f(integral_constant<ci>) through a binary tree of uniform branches (a switch became an LDC jump table + BRX, ~80 cycles).
// The tree tests single BITS of ci, most significant first, so every level's predicate can be formed as soon as ci exists.
template <int BASE, int BIT, int N, class F>
__device__ __forceinline__ void dispatch_bits(uint32_t ci, F &&f) {
    if constexpr (BIT == 0) {
        if constexpr (BASE < N) f(std::integral_constant<int, BASE>{});
    } else {
        constexpr int HALF = BIT >> 1;
        if (ci & (uint32_t)BIT) {
            dispatch_bits<BASE + BIT, HALF, N>(ci, f);
        } else {
            dispatch_bits<BASE, HALF, N>(ci, f);
        }
    }
}
template <int LO, int HI, class F>
__device__ __forceinline__ void dispatch_cell(int ci, F &&f) {
    static_assert(LO == 0 && (HI & (HI - 1)) == 0, "power-of-two cell count");
    dispatch_bits<0, HI / 2, HI>((uint32_t)ci, f);
}

template <int WARPS, int CPW>
struct CellsCfg {
    static constexpr int T = WARPS * 32;
    static constexpr int SLOTS = 4 * CPW;
    static constexpr int NP = WARPS * CPW * kCellPts;
    // main-loop image: X | Y | Z (NP floats each) | ktab (NP u16: inverted ranks) | cell records: 2 buffers x (128 distances | 128 positions) u32
    static constexpr size_t kMain = (size_t)NP * 14 + 2 * 2 * kMaxCells * 4;
    // prepass scratch (aliases the image): hist (4096 int) | ord (NP u16) | red (6 x WARPS float) | wsum (WARPS int)
    static constexpr size_t kPre = (size_t)kOrderCells * 4 + (size_t)NP * 2 + 7 * WARPS * 4;
    static constexpr size_t kSmem = kMain > kPre ? kMain : kPre;
};

// PROF: in-kernel stopwatch (tools/prof_fps_cells.py): per warp, cycles spent in each phase of the round
template <int WARPS, int CPW, bool PROF>
__global__ void __launch_bounds__(WARPS * 32, 1) fps_cells_kernel(const float *__restrict__ xyz, float *__restrict__ temp,
                                                                  int32_t *__restrict__ idx, int n, int m, int log2bs, int cnt,
                                                                  const int32_t *__restrict__ viol, float *__restrict__ new_xyz,
                                                                  unsigned long long *__restrict__ prof) {
    using Cfg = CellsCfg<WARPS, CPW>;
    constexpr int T = Cfg::T, SLOTS = Cfg::SLOTS, NP = Cfg::NP;
    extern __shared__ __align__(16) uint8_t smem[];
    float *X = reinterpret_cast<float *>(smem), *Y = X + NP, *Z = Y + NP;
    uint16_t *ktab = reinterpret_cast<uint16_t *>(Z + NP);
    uint32_t *recs = reinterpret_cast<uint32_t *>(ktab + NP);   // [2 buffers][distance bits | position][kMaxCells]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    xyz += (size_t)cloud * n * 3;
    idx += (size_t)cloud * m;
    if (temp) temp += (size_t)cloud * n;
    if (new_xyz) new_xyz += (size_t)cloud * m * 3;           // optional: coordinates of the picked points, (B, M, 3)
    if (viol != nullptr && __ldg(viol + cloud) == 0) {       // guarded launch, see fps.cu
        for (int i = tid; i < m; i += T) idx[i] = i;
        if (new_xyz)
            for (int i = tid; i < 3 * m; i += T) new_xyz[i] = __ldg(xyz + i);
        return;
    }

    // ------------------------------------------------------------------------------------------------------------
    // prepass 1: Hilbert order of the cloud (counting sort on a 64 x 64 grid over the two widest axes of the bounding box).  The order
    // inside a grid cell is whatever the atomics produce: the sampling result does not depend on which lane owns a point.
    // ------------------------------------------------------------------------------------------------------------
    uint16_t kk[SLOTS];
    {
        int *hist = reinterpret_cast<int *>(smem);
        uint16_t *ord = reinterpret_cast<uint16_t *>(smem + (size_t)kOrderCells * 4);
        float *red = reinterpret_cast<float *>(smem + (size_t)kOrderCells * 4 + (size_t)NP * 2);   // [6][WARPS]
        int *wsum = reinterpret_cast<int *>(red + 6 * WARPS);

        // bounding box of the cloud; the grid spans the two axes with the largest extent (x and z for a LiDAR sweep)
        float lo3[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi3[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (int i = tid; i < n; i += T) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(xyz + (size_t)i * 3 + a);
                lo3[a] = fminf(lo3[a], v); hi3[a] = fmaxf(hi3[a], v);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo3[a] = fminf(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
                hi3[a] = fmaxf(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
            }
            if (lane == 0) { red[(2 * a) * WARPS + warp] = lo3[a]; red[(2 * a + 1) * WARPS + warp] = hi3[a]; }
        }
        for (int i = tid; i < kOrderCells; i += T) hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int w = 0; w < WARPS; ++w) {
                lo3[a] = fminf(lo3[a], red[(2 * a) * WARPS + w]); hi3[a] = fmaxf(hi3[a], red[(2 * a + 1) * WARPS + w]);
            }
        const float ex = hi3[0] - lo3[0], ey = hi3[1] - lo3[1], ez = hi3[2] - lo3[2];
        int ax0 = 0, ax1 = 2;                                  // drop the axis with the smallest extent (ties: keep x, z)
        if (ex < ey && ex <= ez) { ax0 = 1; ax1 = 2; }
        else if (ez < ey && ez < ex) { ax0 = 0; ax1 = 1; }
        const float amin = ax0 == 0 ? lo3[0] : lo3[1], amax = ax0 == 0 ? hi3[0] : hi3[1];
        const float bmin = ax1 == 2 ? lo3[2] : lo3[1], bmax = ax1 == 2 ? hi3[2] : hi3[1];
        const float top = (float)((1 << kOrderBits) - 1);
        const float sa = amax > amin ? top / (amax - amin) : 0.f;
        const float sb = bmax > bmin ? top / (bmax - bmin) : 0.f;
        auto cell_of = [&](int i) -> int {
            const float u = __ldg(xyz + (size_t)i * 3 + ax0), v = __ldg(xyz + (size_t)i * 3 + ax1);
            // non-finite coordinates land in some cell; only the ORDER depends on it
            const float fu = fminf(fmaxf((u - amin) * sa, 0.f), top), fv = fminf(fmaxf((v - bmin) * sb, 0.f), top);
            const uint32_t gu = (uint32_t)(int)fu & ((1u << kOrderBits) - 1u), gv = (uint32_t)(int)fv & ((1u << kOrderBits) - 1u);
            return order_hilbert(gu, gv);
        };
        for (int i = tid; i < n; i += T) atomicAdd(&hist[cell_of(i)], 1);
        __syncthreads();
        constexpr int kPer = kOrderCells / T;
        int v[kPer], run = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { v[j] = hist[tid * kPer + j]; run += v[j]; }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < WARPS ? wsum[lane] : 0;
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            if (lane < WARPS) wsum[lane] = winc - w;
        }
        __syncthreads();
        int excl = wsum[warp] + inc - run;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { hist[tid * kPer + j] = excl; excl += v[j]; }
        __syncthreads();
        for (int i = tid; i < n; i += T) ord[atomicAdd(&hist[cell_of(i)], 1)] = (uint16_t)i;
        __syncthreads();
        // sorted position e -> cell e / 128 -> warp (cell mod WARPS), cell-in-warp (cell / WARPS); slot q of the cell, lane
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int e = (((s >> 2) * WARPS + warp) * kCellPts) + (s & 3) * 32 + lane;
            kk[s] = e < n ? ord[e] : (uint16_t)0xFFFFu;
        }
        __syncthreads();     // the scratch is dead: the image may be written
    }

    // ------------------------------------------------------------------------------------------------------------
    // prepass 2: shared-memory image (coordinates, point indices), running distances, boxes and maxima of my cells
    // ------------------------------------------------------------------------------------------------------------
    float pt[SLOTS];
    float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f, cmax = 0.f;   // cell `lane` (lane < CPW)
    const int cb4 = warp * CPW * 32 + lane;                   // float4 index of (my warp, cell 0, my lane)
    {
        float4 *X4 = reinterpret_cast<float4 *>(X), *Y4 = reinterpret_cast<float4 *>(Y), *Z4 = reinterpret_cast<float4 *>(Z);
        uint2 *K4 = reinterpret_cast<uint2 *>(ktab);
        const float inf = __int_as_float(0x7f800000);
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
            float vx[4], vy[4], vz[4];
            float lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t k = kk[4 * ci + q];
                const bool valid = k != 0xFFFFu;
                vx[q] = 0.f; vy[q] = 0.f; vz[q] = 0.f;
                float t = 0.f;       // padding slot: distance 0 and the worst rank, never wins against a real point
                if (valid) {
                    vx[q] = __ldg(xyz + (size_t)k * 3 + 0);
                    vy[q] = __ldg(xyz + (size_t)k * 3 + 1);
                    vz[q] = __ldg(xyz + (size_t)k * 3 + 2);
                    t = temp ? temp[k] : 1e10f;
                    lo0 = fminf(lo0, vx[q]); hi0 = fmaxf(hi0, vx[q]);
                    lo1 = fminf(lo1, vy[q]); hi1 = fmaxf(hi1, vy[q]);
                    lo2 = fminf(lo2, vz[q]); hi2 = fmaxf(hi2, vz[q]);
                }
                pt[4 * ci + q] = t;
            }
            X4[cb4 + ci * 32] = make_float4(vx[0], vx[1], vx[2], vx[3]);
            Y4[cb4 + ci * 32] = make_float4(vy[0], vy[1], vy[2], vy[3]);
            Z4[cb4 + ci * 32] = make_float4(vz[0], vz[1], vz[2], vz[3]);
            {
                uint32_t ri[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) ri[q] = kk[4 * ci + q] == 0xFFFFu ? 0u : rinv_of(kk[4 * ci + q], log2bs, cnt);
                K4[cb4 + ci * 32] = make_uint2(ri[0] | (ri[1] << 16), ri[2] | (ri[3] << 16));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
                lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
                lo2 = fminf(lo2, __shfl_xor_sync(0xffffffffu, lo2, o)); hi2 = fmaxf(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
            }
            const float lc = fmaxf(fmaxf(pt[4 * ci], pt[4 * ci + 1]), fmaxf(pt[4 * ci + 2], pt[4 * ci + 3]));
            const uint32_t cm = __reduce_max_sync(0xffffffffu, __float_as_uint(lc));
            if (lane == ci) {
                blx = lo0; bly = lo1; blz = lo2; bhx = hi0; bhy = hi1; bhz = hi2;   // an all-padding cell keeps +inf / -inf:
                cmax = __uint_as_float(cm);                                         // lb = +inf, never touched
            }
            // the cell's initial record (both buffers), exactly: with the usual 1e10 start every slot ties
            __syncwarp();
            const uint32_t best = cell_exact_pos(__float_as_uint(lc) == cm, pt[4 * ci], pt[4 * ci + 1], pt[4 * ci + 2], pt[4 * ci + 3],
                                                 lc, (uint32_t)(cb4 * 4 + ci * 128), pn2_smem_u32(ktab));
            if (lane == 0) {
                const int c = warp * CPW + ci;
                recs[c] = cm; recs[kMaxCells + c] = best; recs[2 * kMaxCells + c] = cm; recs[3 * kMaxCells + c] = best;
            }
        }
        for (int i = WARPS * CPW + tid; i < kMaxCells; i += T) {     // entries no warp owns: (0, 0) never beats a real record
            recs[i] = 0u; recs[kMaxCells + i] = 0u; recs[2 * kMaxCells + i] = 0u; recs[3 * kMaxCells + i] = 0u;
        }
    }
    float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);   // idx[0] = 0 always
    if (tid == 0) idx[0] = 0;
    __syncthreads();

    // Shared-memory addresses as 32-bit registers + explicit ld/st.shared: through generic pointers ptxas rebuilt the
    // shared window base (S2UR SR_CgaCtaId -> ULEA) in front of every access group, three dependent ~30-cycle detours
    // per round on the critical path (ncu source page of the first version).
    uint32_t sX = pn2_smem_u32(X);
    asm volatile("" : "+r"(sX));      // opaque: otherwise ptxas re-derives it from SR_CgaCtaId inside the loop
    const uint32_t sY = sX + (uint32_t)NP * 4u, sZ = sX + (uint32_t)NP * 8u, sK = sX + (uint32_t)NP * 12u;
    const uint32_t sRec = sX + (uint32_t)NP * 14u;
    const uint32_t sX4 = sX + (uint32_t)cb4 * 16u, sY4 = sY + (uint32_t)cb4 * 16u, sZ4 = sZ + (uint32_t)cb4 * 16u;
    const uint32_t pos0 = (uint32_t)cb4 * 4u;                 // position of (my warp, cell 0, my lane, slot 0)
    constexpr uint32_t kBufBytes = 2u * kMaxCells * 4u;       // one record buffer: distances | positions
    constexpr uint32_t kPosOff = kMaxCells * 4u;
    const uint32_t rec_mine = sRec + (uint32_t)(warp * CPW) * 4u;            // record of (my warp, cell 0) in buffer 0
    const uint32_t rec_read = sRec + (uint32_t)lane * 16u;                   // the four records lane `lane` reduces

    // One touched cell: update its 4 running distances per lane, new cell maximum, and the cell's RECORD (maximum,
    // position of the point that holds it) for this round.  Common case: one lane and one slot hold the maximum -> that
    // lane writes the record; any tie takes the exact path (reference rank).
    uint32_t cpos = 0u;      // lane ci < CPW: position in the record of my cell ci (the record's distance is cmax)
    auto update_cell = [&](auto CI, const float4 x4, const float4 y4, const float4 z4, const float2 ncx, const float2 ncy,
                           const float2 ncz, const uint32_t rec_w) {
        constexpr int ci = decltype(CI)::value;
        // (p - c) == p + (-c) exactly; t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t) as pn2_sqdist
        const float2 dx0 = __fadd2_rn(make_float2(x4.x, x4.y), ncx), dx1 = __fadd2_rn(make_float2(x4.z, x4.w), ncx);
        const float2 dy0 = __fadd2_rn(make_float2(y4.x, y4.y), ncy), dy1 = __fadd2_rn(make_float2(y4.z, y4.w), ncy);
        const float2 dz0 = __fadd2_rn(make_float2(z4.x, z4.y), ncz), dz1 = __fadd2_rn(make_float2(z4.z, z4.w), ncz);
        float2 t0 = __fmul2_rn(dy0, dy0), t1 = __fmul2_rn(dy1, dy1);
        t0 = __ffma2_rn(dx0, dx0, t0); t1 = __ffma2_rn(dx1, dx1, t1);
        t0 = __ffma2_rn(dz0, dz0, t0); t1 = __ffma2_rn(dz1, dz1, t1);
        const float p0 = fminf(t0.x, pt[4 * ci + 0]), p1 = fminf(t0.y, pt[4 * ci + 1]);
        const float p2 = fminf(t1.x, pt[4 * ci + 2]), p3 = fminf(t1.y, pt[4 * ci + 3]);
        pt[4 * ci + 0] = p0; pt[4 * ci + 1] = p1; pt[4 * ci + 2] = p2; pt[4 * ci + 3] = p3;
        const float mm = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
        const uint32_t cm = __reduce_max_sync(0xffffffffu, __float_as_uint(mm));
        const bool e0 = p0 == mm, e1 = p1 == mm, e2 = p2 == mm, e3 = p3 == mm;
        const uint32_t q = e0 ? 0u : (e1 ? 1u : (e2 ? 2u : 3u));
        const bool several = ((int)e0 + (int)e1 + (int)e2 + (int)e3) > 1;
        const bool cand = __float_as_uint(mm) == cm;
        const uint32_t cbal = __ballot_sync(0xffffffffu, cand);
        const uint32_t amb = __ballot_sync(0xffffffffu, cand && several);
        const uint32_t posc = pos0 + (uint32_t)(ci * 128);
        uint32_t wpos;
        if (amb == 0u && (cbal & (cbal - 1u)) == 0u) {
            sts_u32_if(cand, rec_w + ci * 4, cm);                        // predicated stores: no divergent branch
            sts_u32_if(cand, rec_w + ci * 4 + kPosOff, posc + q);
            wpos = __reduce_max_sync(0xffffffffu, cand ? posc + q : 0u);     // for lane ci; needed only after the barrier
        } else {
            wpos = cell_exact_pos(cand, p0, p1, p2, p3, mm, posc, sK);
            if (lane == 0) { sts_u32(rec_w + ci * 4, cm); sts_u32(rec_w + ci * 4 + kPosOff, wpos); }
        }
        if (lane == ci) { cmax = __uint_as_float(cm); cpos = wpos; }
    };

    uint32_t kpend = 0u;
    unsigned long long p_test = 0, p_upd = 0, p_rec = 0, p_bar_u = 0, p_bar_n = 0, p_red = 0, p_nupd = 0, p_cells = 0;
    for (int r = 0; r < m - 1; ++r) {
        const uint32_t boff = (uint32_t)(r & 1) * kBufBytes;
        const long long c0 = PROF ? clock64() : 0;
        // 1. which of my cells can the new centre change?  (same float expression as the distance itself: see the header)
        const float bx = fmaxf(fmaxf(__fadd_rn(blx, -cx), __fadd_rn(cx, -bhx)), 0.f);
        const float by = fmaxf(fmaxf(__fadd_rn(bly, -cy), __fadd_rn(cy, -bhy)), 0.f);
        const float bz = fmaxf(fmaxf(__fadd_rn(blz, -cz), __fadd_rn(cz, -bhz)), 0.f);
        const float lb = pn2_sqdist(bx, by, bz);
        if (tid == 0 && new_xyz) {      // pick r (its coordinates were just consumed above: no extra scoreboard wait)
            new_xyz[3 * r] = cx; new_xyz[3 * r + 1] = cy; new_xyz[3 * r + 2] = cz;
        }
        const bool touched = lane < CPW && lb < cmax;
        const uint32_t mask = __ballot_sync(0xffffffffu, touched);
        int ci = (int)__reduce_min_sync(0xffffffffu, touched ? (uint32_t)lane : 32u);   // first touched cell (redux: ~15 cycles, BREV + FLO of the mask: ~40)
        const long long c1 = PROF ? clock64() : 0;
        if (mask) {
            // 2. update the touched cells (usually one) and their records.  The cell's coordinates are requested before the
            // dispatch (dynamic address), the branch tree that selects the registers of cell ci runs under the load latency.
            const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
            uint32_t todo = mask;
            do {
                const float4 x4 = lds_f4(sX4 + ci * 512), y4 = lds_f4(sY4 + ci * 512), z4 = lds_f4(sZ4 + ci * 512);
                todo &= ~(1u << ci);
                dispatch_cell<0, CPW>(ci, [&](auto CI) { update_cell(CI, x4, y4, z4, ncx, ncy, ncz, rec_mine + boff); });
                ci = __ffs(todo) - 1;
            } while (todo);
            if (PROF) { p_upd += clock64() - c1; p_nupd += 1; p_cells += __popc(mask); }
        }
        const long long c4 = PROF ? clock64() : 0;
        __syncthreads();
        // 3. every warp reduces the 128 cell records (redundantly, four per lane) -> next centre
        const uint4 h4 = lds_u4(rec_read + boff), p4 = lds_u4(rec_read + boff + kPosOff);
        const uint32_t m01 = h4.x > h4.y ? h4.x : h4.y, m23 = h4.z > h4.w ? h4.z : h4.w;
        const uint32_t lm = m01 > m23 ? m01 : m23;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, lm);
        long long c5 = 0;
        if (PROF) { c5 = clock64() + (mh == 0x12345678u ? 1 : 0); }
        const bool f0 = h4.x == mh, f1 = h4.y == mh, f2 = h4.z == mh, f3 = h4.w == mh;
        const bool mine = f0 || f1 || f2 || f3;
        const bool several = ((int)f0 + (int)f1 + (int)f2 + (int)f3) > 1;
        const uint32_t who = __ballot_sync(0xffffffffu, mine);
        const uint32_t amb = __ballot_sync(0xffffffffu, mine && several);
        // one candidate: its position reaches every lane through a second redux (shorter than FLO + SHFL of the ballot)
        uint32_t pos = __reduce_max_sync(0xffffffffu, mine ? (f0 ? p4.x : (f1 ? p4.y : (f2 ? p4.z : p4.w))) : 0u);
        if (amb != 0u || (who & (who - 1u)) != 0u)
            pos = records_exact_pos(h4, p4, mh, sK);      // several cells tie: the reference rank decides
        cx = lds_f32(sX + pos * 4u); cy = lds_f32(sY + pos * 4u); cz = lds_f32(sZ + pos * 4u);
        // idx[r + 1]: the table look-up is issued now, the global store waits until the end of the next round (a store right
        // here stalled every warp on the look-up's scoreboard, predicated off or not)
        if (tid == 0) {
            if (r > 0) idx[r] = (int32_t)kpend;      // the table entry (inverted rank); turned into point indices after the loop
            kpend = lds_u16(sK + pos * 2u);
        }
        // the records written in this round also go into the other buffer (read in the next round): only now is nobody
        // reading it any more -- every warp has passed this round's barrier
        if (touched) {
            const uint32_t o = rec_mine + (boff ^ kBufBytes) + (uint32_t)lane * 4u;
            sts_u32(o, __float_as_uint(cmax));
            sts_u32(o + kPosOff, cpos);
        }
        __syncwarp();      // orders these stores before the record another lane of this warp may write to the same slot next round
        if (PROF) {
            const long long c6 = clock64() + (cz == 1.2345e-30f ? 1 : 0);
            p_test += c1 - c0;
            if (mask) p_bar_u += c5 - c4; else p_bar_n += c5 - c4;
            p_red += c6 - c5;
        }
    }
    if (tid == 0 && m > 1) idx[m - 1] = (int32_t)kpend;
    // idx[1 .. m-1] hold inverted ranks (the division that recovers the point index would sit in every warp's round, predicated
    // off or not): converted here by the whole CTA
    __syncthreads();
    for (int i = 1 + tid; i < m; i += T) idx[i] = (int32_t)k_of_rinv((uint32_t)idx[i], log2bs, cnt);
    if (tid == 0 && new_xyz) { new_xyz[3 * (m - 1)] = cx; new_xyz[3 * (m - 1) + 1] = cy; new_xyz[3 * (m - 1) + 2] = cz; }
    if (PROF && lane == 0 && prof) {
        unsigned long long *o = prof + ((size_t)cloud * WARPS + warp) * 8;
        o[0] = p_test; o[1] = p_upd; o[2] = p_rec; o[3] = p_bar_u; o[4] = p_bar_n; o[5] = p_red; o[6] = p_nupd; o[7] = p_cells;
    }

    if (temp) {   // the reference leaves the running min distances in the caller's scratch
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const uint32_t ri = ktab[pos0 + (uint32_t)((s >> 2) * 128 + (s & 3))];
            if (ri != 0u) temp[k_of_rinv(ri, log2bs, cnt)] = pt[s];
        }
    }
}

unsigned long long *g_cells_prof = nullptr;

template <int WARPS, int CPW>
cudaError_t launch_cells(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int log2bs, int cnt,
                         const int32_t *viol, float *new_xyz, cudaStream_t stream) {
    using Cfg = CellsCfg<WARPS, CPW>;
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        cudaError_t ea = cudaFuncSetAttribute(fps_cells_kernel<WARPS, CPW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)Cfg::kSmem);
        if constexpr (WARPS * CPW == 128) {
            if (ea == cudaSuccess)
                ea = cudaFuncSetAttribute(fps_cells_kernel<WARPS, CPW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        }
        if (ea != cudaSuccess) return ea;
        attr_done = true;
    }
    if constexpr (WARPS * CPW == 128) {      // stopwatch build: the 16384-point shapes only (tools/prof_fps_cells.py)
        if (g_cells_prof) {
            fps_cells_kernel<WARPS, CPW, true><<<b, Cfg::T, Cfg::kSmem, stream>>>(xyz, temp, idx, n, m, log2bs, cnt, viol, new_xyz, g_cells_prof);
            return cudaGetLastError();
        }
    }
    fps_cells_kernel<WARPS, CPW, false><<<b, Cfg::T, Cfg::kSmem, stream>>>(xyz, temp, idx, n, m, log2bs, cnt, viol, new_xyz, nullptr);
    return cudaGetLastError();
}

}  // namespace

extern "C" int pn2_fps_ref_block_size(int n);

// stopwatch buffer for tools/prof_fps_cells.py: 8 u64 per (cloud, warp) in device memory, or NULL (never used by the product)
PN2_API void pn2_fps_cells_set_profile(void *buf) { g_cells_prof = static_cast<unsigned long long *>(buf); }

// Does the cell kernel take clouds of n points?  (one CTA's shared memory holds 16384 points)
bool pn2_fps_cells_supported(int n) { return n >= 1 && n <= 16384; }

// Internal door for fps.cu (pn2_fps_f32's heuristic) and pn2_fps_cells_f32.  warps: 0 = heuristic, or 4 / 8 / 16 to force
// the number of warps of the CTA (tests, tuning); returns cudaErrorInvalidValue for a combination that is not built.
cudaError_t pn2_fps_cells_launch(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int warps,
                                 const int32_t *viol, float *new_xyz, cudaStream_t stream) {
    const int bs = pn2_fps_ref_block_size(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    const int cnt = (n + bs - 1) / bs;
    const int cells = (n + kCellPts - 1) / kCellPts;      // 1 .. 128
    if (warps == 0) warps = 8;       // measured best at every size (tools/bench_fps_cluster.py): the redundant per-warp work of a round grows with the warp count
#define PN2_CELLS_GO(W, C) return launch_cells<W, C>(xyz, temp, idx, b, n, m, log2bs, cnt, viol, new_xyz, stream)
    if (warps == 4) {
        if (cells <= 32) PN2_CELLS_GO(4, 8);
        if (cells <= 64) PN2_CELLS_GO(4, 16);
        if (cells <= 128) PN2_CELLS_GO(4, 32);
    } else if (warps == 8) {
        if (cells <= 8) PN2_CELLS_GO(8, 1);
        if (cells <= 16) PN2_CELLS_GO(8, 2);
        if (cells <= 32) PN2_CELLS_GO(8, 4);
        if (cells <= 64) PN2_CELLS_GO(8, 8);
        if (cells <= 128) PN2_CELLS_GO(8, 16);
    } else if (warps == 16) {
        if (cells <= 16) PN2_CELLS_GO(16, 1);
        if (cells <= 32) PN2_CELLS_GO(16, 2);
        if (cells <= 64) PN2_CELLS_GO(16, 4);
        if (cells <= 128) PN2_CELLS_GO(16, 8);
    }
#undef PN2_CELLS_GO
    return cudaErrorInvalidValue;
}

// pn2_fps_f32 through the pruned one-CTA kernel, whatever the heuristic of pn2_fps_f32 would pick (tests, tuning).
// n <= 16384; warps = 0 (heuristic), 4, 8 or 16.  Same result as pn2_fps_f32 for every value.
PN2_API int pn2_fps_cells_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int warps,
                              cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || (!xyz && b * n > 0) || (!idx && b * m > 0)) {
        pn2_set_last_error("pn2_fps_cells_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || m == 0) return PN2_OK;
    if (!pn2_fps_cells_supported(n)) {
        pn2_set_last_error("pn2_fps_cells_f32: 1 <= N <= 16384 points per cloud");
        return PN2_ERR_UNSUPPORTED;
    }
    const cudaError_t e = pn2_fps_cells_launch(xyz, temp, idx, b, n, m, warps, nullptr, nullptr, stream);
    if (e == cudaErrorInvalidValue) {
        pn2_set_last_error("pn2_fps_cells_f32: warps must be 0, 4, 8 or 16");
        return PN2_ERR_INVALID;
    }
    if (e != cudaSuccess) {
        pn2_set_last_error(cudaGetErrorString(e));
        return PN2_ERR_LAUNCH;
    }
    return PN2_OK;
}
