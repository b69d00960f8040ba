// tc_producer.cuh -- the operand producer shared by the tcgen05 shared-MLP kernels.
//
// Sixteen producer warps turn fp32 activation rows in global memory (plain rows, or rows gathered
// through ball-query indices with the xyz half of the first SA layer and its ReLU applied on the
// fly) into the bf16 hi/lo, 128-byte-swizzled K-major tiles the MMA reads from shared memory.
//
// The design target is memory-level parallelism: a 128 x 64 K-block is 32 KB of fp32 and the
// tensor pipe consumes one every ~770 cycles, against ~1500 cycles of L2/HBM latency.  Two groups
// of eight warps fill two ring stages concurrently, each thread with 8 x 16 B of loads in
// flight in registers (the register file, otherwise idle in a TMEM-accumulator kernel, is the
// landing zone: 64 KB outstanding per SM) and no shared-memory staging.
// Row metadata for gathered tiles (source row, xyz offset to the centre) is produced by a
// separate meta warp running up to four tiles ahead through its own mbarrier ring, so the
// dependent idx -> xyz loads never sit on the producers' critical path.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int kBM = 128;               // rows per tile (UMMA M)
constexpr int kBK = 64;                // bf16 per K-block row (one 128-byte swizzle atom)
constexpr int kTileBytes = kBM * kBK * 2;   // one bf16 operand tile (hi or lo): 16 KB
constexpr int kProdWarps = 16;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMetaDepth = 4;          // tiles of row metadata in flight

struct RowMeta {
    int src;                           // source row of x (rows beyond the end point at row 0; the epilogue masks them)
    float dx, dy, dz;
};

struct ProducerArgs {
    const float *x; int ldx; int cin; long long rows; int vec_ok;
    const float *x2; int ldx2; int kb_split;   // K-blocks >= kb_split come from x2 (plain rows, FAST path only)
    const int32_t *idx; const float *xyz; const float *centres; int n, m, ns;   // gather only
    int nkb, stages, nchunks;
    long long items;                   // work items of the whole grid; item -> tile = item / nchunks
    uint8_t *ring; uint32_t stage_bytes;   // A.hi at ring + s * stage_bytes, A.lo kTileBytes after it
    uint64_t *full, *empty;            // per stage
    RowMeta *meta; uint64_t *meta_full, *meta_empty;   // [kMetaDepth] x kBM, gather only
    const float *wxs; int kpad;        // (3, kpad) xyz coefficients of layer 1 in shared memory
    unsigned long long *prof;          // optional stopwatch buffer (32 u64 per CTA, tools/prof_tc.py) or nullptr
};

// volatile 16-byte shared load (so the compiler re-reads instead of pinning 12 registers)
__device__ __forceinline__ void lds128(const float *p, float4 &v) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(pn2_smem_u32(p)));
}

// ---- meta warp (gather only): RowMeta of tile `it` into slot it % kMetaDepth ----
// One warp feeds the sixteen producers, so its two dependent global round trips per tile
// (idx -> xyz[idx]) are kept off the critical path: the four idx loads of a lane are independent and
// issued together, the idx loads of the NEXT tile are issued before this tile's coordinates are
// consumed, and rows past the end are clamped (not branched around) so nothing serialises the loads.
// Row -> (centre, cloud) uses one 64-bit division per TILE (uniform) and 32-bit ones per row.
template <bool PROF>
__device__ __forceinline__ void meta_run(const ProducerArgs &a, int lane) {
    const long long first = blockIdx.x, stride = gridDim.x;
    const uint32_t ns = (uint32_t)a.ns, m = (uint32_t)a.m;
    const long long last_row = a.rows - 1;
    constexpr int Q = kBM / 32;
    unsigned long long w_slot = 0;
    const long long t_begin = PROF ? clock64() : 0;
    int jn[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) jn[q] = 0;
    if (first < a.items) {
        const long long row0 = (first / a.nchunks) * kBM;
#pragma unroll
        for (int q = 0; q < Q; ++q) jn[q] = __ldg(a.idx + min(row0 + q * 32 + lane, last_row));
    }
    long long it = 0;
    for (long long item = first; item < a.items; item += stride, ++it) {
        const int slot = (int)(it % kMetaDepth);
        const long long row0 = (item / a.nchunks) * kBM;
        const long long centre0 = row0 / ns;                       // tile-uniform 64-bit part
        const uint32_t rem0 = (uint32_t)(row0 - centre0 * ns);
        const long long cloud0 = centre0 / m;
        const uint32_t crem0 = (uint32_t)(centre0 - cloud0 * m);
        int src[Q];
        float pj[Q][3], pc[Q][3];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const uint32_t r = (uint32_t)(q * 32 + lane);
            const bool live = row0 + r <= last_row;
            const uint32_t rr = live ? r : (uint32_t)(last_row - row0);   // clamp to the last valid row of the tile
            const uint32_t dc = (rem0 + rr) / ns;
            const long long centre = centre0 + dc;
            const long long cloud = cloud0 + (crem0 + dc) / m;
            const long long s = cloud * a.n + jn[q];
            src[q] = live ? (int)s : 0;          // rows beyond the end: any valid row (masked by the epilogue)
            const float *gj = a.xyz + s * 3, *gc = a.centres + centre * 3;
            pj[q][0] = __ldg(gj); pj[q][1] = __ldg(gj + 1); pj[q][2] = __ldg(gj + 2);
            pc[q][0] = __ldg(gc); pc[q][1] = __ldg(gc + 1); pc[q][2] = __ldg(gc + 2);
            if (!live) { pj[q][0] = pc[q][0]; pj[q][1] = pc[q][1]; pj[q][2] = pc[q][2]; }
        }
        // next tile's neighbour indices: in flight while this tile's coordinates arrive
        if (item + stride < a.items) {
            const long long nrow0 = ((item + stride) / a.nchunks) * kBM;
#pragma unroll
            for (int q = 0; q < Q; ++q) jn[q] = __ldg(a.idx + min(nrow0 + q * 32 + lane, last_row));
        }
        mbar_wait_timed<PROF>(&a.meta_empty[slot], (uint32_t)((it / kMetaDepth) & 1) ^ 1, w_slot);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            RowMeta mt;
            mt.src = src[q];
            mt.dx = pj[q][0] - pc[q][0];
            mt.dy = pj[q][1] - pc[q][1];
            mt.dz = pj[q][2] - pc[q][2];
            a.meta[slot * kBM + q * 32 + lane] = mt;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a.meta_full[slot]);
    }
    if (PROF && lane == 0) {
        unsigned long long *o = a.prof + (size_t)blockIdx.x * 32;
        o[17] = w_slot;
        o[18] = (unsigned long long)(clock64() - t_begin);
    }
}

// ---- producer warps.  hook(item, kb, stage) runs on the first thread of the group that owns the
//      step once the stage is free (linear_tc.cu launches the weight K-block's bulk copy there).
//
// The sixteen warps form kGroups = 2 groups that take alternate K-block steps, so two ring stages
// are being filled concurrently and a warp has exactly ONE batch of loads (8 x 16 B per lane) in
// flight: deeper per-warp queues do not work, a warp has six scoreboard slots and later short
// loads end up waiting on the slot of the oldest global batch (profiled: long_scoreboard stalls
// on the shared-memory reads that follow the prefetch).  A group issues the loads of its next step
// right after storing the current one, i.e. before it waits for that step's stage to drain, so
// the flight time overlaps the MMA of the other group's stage. ----
constexpr int kGroups = 2;
constexpr int kGroupWarps = kProdWarps / kGroups;          // 8
constexpr int kRowsPerPass = 2 * kGroupWarps;               // 16
constexpr int kPasses = kBM / kRowsPerPass;                 // 8 float4 per thread per K-block

template <bool GATHER, bool FAST, bool PROF, class StageHook>
__device__ __forceinline__ void producer_body(const ProducerArgs &a, int ptid, StageHook hook) {
    const int lane = ptid & 31, pw = ptid >> 5;
    const int group = pw % kGroups, wg = pw / kGroups;
    const int rsub = wg * 2 + (lane >> 4);     // row inside a 16-row pass: row r = ps * 16 + rsub
    const int kq = (lane & 15) * 4;            // first k of this thread's float4 inside a K-block
    // byte offset of this thread's 8 bytes inside a swizzled tile, pass 0; pass ps adds ps * 2048
    // (r >> 3 = 2 ps + (rsub >> 3), r & 7 = rsub & 7: the swizzle term does not depend on ps)
    const uint32_t toff = (uint32_t)(((rsub >> 3) << 10) + ((rsub & 7) << 7)) +
                          ((((uint32_t)((lane & 15) >> 1) ^ (uint32_t)(rsub & 7)) << 4) | ((uint32_t)(lane & 1) << 3));
    const long long first = blockIdx.x, stride = gridDim.x;
    // per-CTA counts fit 32 bits (a CTA walks at most items / gridDim.x tiles): 64-bit bookkeeping in this
    // loop was a measurable share of the producers' instructions
    const int my_items = first < a.items ? (int)((a.items - first + stride - 1) / stride) : 0;
    const int total_steps = my_items * a.nkb;

    // position of the step whose loads are issued next (l_*) and of the step stored next (s_*)
    int l_t = group, s_t = group;
    int l_it = group / a.nkb, s_it = l_it;
    int l_kb = group % a.nkb, s_kb = l_kb;
    int meta_seen = -1, meta_freed = 0;   // tiles whose meta this warp has waited for / released
    int stage = group % a.stages;
    uint32_t phase = (uint32_t)((group / a.stages) & 1);

    unsigned long long w_meta = 0, w_stage = 0;   // stopwatch (a.prof): cycles blocked on row metadata / on a free stage
    const long long t_begin = PROF ? clock64() : 0;
    float4 areg[kPasses];
    auto issue_loads = [&]() {
        const int k = l_kb * kBK + kq;
        const RowMeta *mt = a.meta + (l_it % kMetaDepth) * kBM + rsub;
        long long row0 = 0;
        if (GATHER) {
            if (l_it != meta_seen) {
                mbar_wait_timed<PROF>(&a.meta_full[l_it % kMetaDepth], (uint32_t)((l_it / kMetaDepth) & 1), w_meta);
                meta_seen = l_it;
            }
        } else {
            row0 = ((first + l_it * stride) / a.nchunks) * kBM + rsub;
        }
        const float *xk = a.x + k;
        if (FAST) {
            long long ld = a.ldx;
            if (!GATHER && l_kb >= a.kb_split) {   // second operand source (concatenated input)
                xk = a.x2 + (k - a.kb_split * kBK);
                ld = a.ldx2;
            }
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
                long long src;
                if (GATHER) src = mt[ps * kRowsPerPass].src;
                else src = min(row0 + ps * kRowsPerPass, a.rows - 1);
                areg[ps] = __ldg(reinterpret_cast<const float4 *>(xk + src * ld));
            }
        } else {
            const bool kvec = a.vec_ok && k + 3 < a.cin;
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
                long long src;
                if (GATHER) src = mt[ps * kRowsPerPass].src;
                else src = min(row0 + ps * kRowsPerPass, a.rows - 1);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < a.cin) {
                    const float *px = xk + src * a.ldx;
                    if (kvec) {
                        v = __ldg(reinterpret_cast<const float4 *>(px));
                    } else {
                        v.x = __ldg(px);
                        if (k + 1 < a.cin) v.y = __ldg(px + 1);
                        if (k + 2 < a.cin) v.z = __ldg(px + 2);
                        if (k + 3 < a.cin) v.w = __ldg(px + 3);
                    }
                }
                areg[ps] = v;
            }
        }
        l_t += kGroups;
        l_kb += kGroups;
        while (l_kb >= a.nkb) { l_kb -= a.nkb; ++l_it; }
    };

    if (l_t < total_steps) issue_loads();
    while (s_t < total_steps) {
        uint8_t *sbase = a.ring + (size_t)stage * a.stage_bytes + toff;
        mbar_wait_lazy_timed<PROF>(&a.empty[stage], phase ^ 1, w_stage);
        if (wg == 0 && lane == 0) hook(first + s_it * stride, s_kb, stage);
        const int k = s_kb * kBK + kq;
        const RowMeta *mt = a.meta + (s_it % kMetaDepth) * kBM + rsub;
        float4 w0, w1, w2;
        if (GATHER) {
            w0 = *reinterpret_cast<const float4 *>(a.wxs + k);
            w1 = *reinterpret_cast<const float4 *>(a.wxs + a.kpad + k);
            w2 = *reinterpret_cast<const float4 *>(a.wxs + 2 * a.kpad + k);
        }
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) {
            float4 v = areg[ps];
            if (GATHER) {
                // layer-1 output of this (centre, neighbour) pair: relu(H_j + W1x . (x_j - centre))
                const float4 q = *reinterpret_cast<const float4 *>(mt + ps * kRowsPerPass);   // (src bits, dx, dy, dz)
                // packed fp32 FMAs (two channels per instruction; each half rounds like fmaf)
                const float2 qy = make_float2(q.y, q.y), qz = make_float2(q.z, q.z), qw = make_float2(q.w, q.w);
                float2 t01 = __ffma2_rn(make_float2(w0.x, w0.y), qy, make_float2(v.x, v.y));
                float2 t23 = __ffma2_rn(make_float2(w0.z, w0.w), qy, make_float2(v.z, v.w));
                t01 = __ffma2_rn(make_float2(w1.x, w1.y), qz, t01);
                t23 = __ffma2_rn(make_float2(w1.z, w1.w), qz, t23);
                t01 = __ffma2_rn(make_float2(w2.x, w2.y), qw, t01);
                t23 = __ffma2_rn(make_float2(w2.z, w2.w), qw, t23);
                v.x = fmaxf(t01.x, 0.f);
                v.y = fmaxf(t01.y, 0.f);
                v.z = fmaxf(t23.x, 0.f);
                v.w = fmaxf(t23.y, 0.f);
                if (!FAST && k + 3 >= a.cin) {     // keep the K padding at zero
                    if (k + 0 >= a.cin) v.x = 0.f;
                    if (k + 1 >= a.cin) v.y = 0.f;
                    if (k + 2 >= a.cin) v.z = 0.f;
                    v.w = 0.f;
                }
            }
            uint2 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<uint2 *>(sbase + ps * 2048) = hi;
            *reinterpret_cast<uint2 *>(sbase + kTileBytes + ps * 2048) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a.full[stage]);
        // advance to this group's next step; release the metadata of the tiles left behind
        s_t += kGroups;
        s_kb += kGroups;
        while (s_kb >= a.nkb) { s_kb -= a.nkb; ++s_it; }
        if (GATHER && lane == 0) {
            const int upto = s_t < total_steps ? s_it : my_items;
            for (; meta_freed < upto; ++meta_freed) mbar_arrive(&a.meta_empty[meta_freed % kMetaDepth]);
        }
        stage += kGroups;
        while (stage >= a.stages) { stage -= a.stages; phase ^= 1; }
        if (l_t < total_steps) issue_loads();
    }
    if (PROF && ptid == 0) {
        unsigned long long *o = a.prof + (size_t)blockIdx.x * 32;
        o[14] = w_meta;
        o[15] = w_stage;
        o[16] = (unsigned long long)(clock64() - t_begin);
    }
}

// ---- producers of a layer whose INPUT is itself a tiny first layer computed on the fly:
//        A[row][k] = relu(bpre[k] + sum_{c < CPRE} Wpre[c][k] * x[row][c])          (fp32 FMAs, packed)
//      e.g. RCNN xyz_up_layer (rcnn_net.py:41-47: SharedMLP [5 -> 128 -> 128]): the 128-channel output of its first
//      layer (0.42 GB written and read back per step) never exists; the producers read 5 floats per row instead of 128.
//      a.wxs holds (CPRE + 1, kpad) floats: the CPRE weight rows, then the bias row.  Plain consecutive rows, nchunks = 1,
//      rows of x 16-byte aligned.  No global prefetch structure is needed: the inputs are 20 bytes per row.
template <int CPRE, class StageHook>
__device__ __forceinline__ void producer_pre(const ProducerArgs &a, int ptid, StageHook hook) {
    static_assert(CPRE >= 1 && CPRE <= 8, "first layer of up to 8 input channels");
    const int lane = ptid & 31, pw = ptid >> 5;
    const int group = pw % kGroups, wg = pw / kGroups;
    const int rsub = wg * 2 + (lane >> 4);
    const int kq = (lane & 15) * 4;
    const uint32_t toff = (uint32_t)(((rsub >> 3) << 10) + ((rsub & 7) << 7)) +
                          ((((uint32_t)((lane & 15) >> 1) ^ (uint32_t)(rsub & 7)) << 4) | ((uint32_t)(lane & 1) << 3));
    const long long first = blockIdx.x, stride = gridDim.x;
    const int my_items = first < a.items ? (int)((a.items - first + stride - 1) / stride) : 0;
    const int total_steps = my_items * a.nkb;
    int s_t = group, s_it = group / a.nkb, s_kb = group % a.nkb;
    int stage = group % a.stages;
    uint32_t phase = (uint32_t)((group / a.stages) & 1);
    while (s_t < total_steps) {
        uint8_t *sbase = a.ring + (size_t)stage * a.stage_bytes + toff;
        mbar_wait_lazy(&a.empty[stage], phase ^ 1);
        if (wg == 0 && lane == 0) hook(first + (long long)s_it * stride, s_kb, stage);
        const int k = s_kb * kBK + kq;
        float2 w01[CPRE + 1], w23[CPRE + 1];               // weights of this thread's 4 channels; row CPRE = bias
#pragma unroll
        for (int c = 0; c <= CPRE; ++c) {
            const float4 w = *reinterpret_cast<const float4 *>(a.wxs + c * a.kpad + k);
            w01[c] = make_float2(w.x, w.y);
            w23[c] = make_float2(w.z, w.w);
        }
        const long long row0 = (first + (long long)s_it * stride) * kBM + rsub;
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) {
            const long long row = min(row0 + ps * kRowsPerPass, a.rows - 1);
            const float *xr = a.x + row * a.ldx;
            float in[8];
            const float4 q0 = __ldg(reinterpret_cast<const float4 *>(xr));
            in[0] = q0.x; in[1] = q0.y; in[2] = q0.z; in[3] = q0.w;
            if (CPRE > 4) {
                const float4 q1 = __ldg(reinterpret_cast<const float4 *>(xr + 4));   // the row has >= 8 floats (checked on the host)
                in[4] = q1.x; in[5] = q1.y; in[6] = q1.z; in[7] = q1.w;
            }
            float2 t01 = w01[CPRE], t23 = w23[CPRE];
#pragma unroll
            for (int c = 0; c < CPRE; ++c) {
                const float2 xc = make_float2(in[c], in[c]);
                t01 = __ffma2_rn(w01[c], xc, t01);
                t23 = __ffma2_rn(w23[c], xc, t23);
            }
            float4 v = make_float4(fmaxf(t01.x, 0.f), fmaxf(t01.y, 0.f), fmaxf(t23.x, 0.f), fmaxf(t23.y, 0.f));
            if (k + 3 >= a.cin) {                          // keep the K padding at zero
                if (k + 0 >= a.cin) v.x = 0.f;
                if (k + 1 >= a.cin) v.y = 0.f;
                if (k + 2 >= a.cin) v.z = 0.f;
                if (k + 3 >= a.cin) v.w = 0.f;
            }
            uint2 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<uint2 *>(sbase + ps * 2048) = hi;
            *reinterpret_cast<uint2 *>(sbase + kTileBytes + ps * 2048) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a.full[stage]);
        s_t += kGroups;
        s_kb += kGroups;
        while (s_kb >= a.nkb) { s_kb -= a.nkb; ++s_it; }
        stage += kGroups;
        while (stage >= a.stages) { stage -= a.stages; phase ^= 1; }
    }
}

// FAST: 16-byte aligned rows and K a multiple of 64 -- no per-element bounds in the inner loops.  The
// choice is a KERNEL template parameter (made on the host by producer_fast()): with both bodies in one
// kernel the producers' code doubles and the roles evict each other from the instruction cache.
inline bool producer_fast(int vec_ok, int cin) { return vec_ok && (cin & (kBK - 1)) == 0; }

template <bool GATHER, bool FAST, bool PROF, class StageHook>
__device__ __forceinline__ void producer_run(const ProducerArgs &a, int ptid, StageHook hook) {
    producer_body<GATHER, FAST, PROF>(a, ptid, hook);
}

}  // namespace tc
