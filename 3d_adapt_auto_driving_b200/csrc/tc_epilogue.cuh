// tc_epilogue.cuh -- epilogue pieces shared by the tcgen05 shared-MLP kernels.
//
// Eight epilogue warps: warp w reads TMEM lane quadrant q = w & 3 (rows 32q .. 32q+31 of the
// tile, one row per lane -- a hardware restriction of tcgen05.ld) and the 16-column chunks with
// index = (w >> 2) mod 2, so two warps share every quadrant and split its columns.  With four
// warps the epilogue, not the tensor pipe or the producers, bounded the kernels (ncu: 60 % of
// the epilogue warps' samples were in the chunk loops at 0.26 IPC); the two warps of a scheduler
// overlap each other's TMEM load latency.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int kBM_ = 128;           // rows per tile
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;

// ---- pooled epilogue: max over the rows of each pooling group of ReLU(acc + bias) ----
// relu(. + b) is monotone, so the max runs on the raw accumulators and bias / ReLU are applied to
// the one surviving value per lane.  The cross-lane maximum is a butterfly that exchanges half of
// the remaining values at every step (16 columns over 32 lanes: 8 + 4 + 2 + 1 + 1 shuffles), after
// which lane l holds column (l >> 1) of the chunk (pool >= 32) or column (l & 15) of its half-warp's
// group (pool 16).  redux.sync was tried first: ~60 cycles each through the uniform datapath, 16
// per chunk, 4.7 k cycles per tile.
__device__ __forceinline__ float bfly_max_pair(float keep, float send, int mask) {
    return fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, mask));
}

// Cross-lane maximum of one 16-column chunk (raw accumulators, one row per lane).  Returns, in lane l,
// column (l >> 1) of the chunk (pool >= 32: all 32 rows are one group) or column (l & 15) of the
// half-warp's group (pool 16).
__device__ __forceinline__ float chunk_max(const uint32_t (&raw)[16], bool live, bool all_live, bool pool16, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
    if (!all_live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = live ? v[j] : -INFINITY;
    }
    float w8[8], w4[4], w2[2], r;
    if (pool16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = bfly_max_pair(b3 ? v[i + 8] : v[i], b3 ? v[i] : v[i + 8], 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) w4[i] = bfly_max_pair(b2 ? w8[i + 4] : w8[i], b2 ? w8[i] : w8[i + 4], 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) w2[i] = bfly_max_pair(b1 ? w4[i + 2] : w4[i], b1 ? w4[i] : w4[i + 2], 2);
        r = bfly_max_pair(b0 ? w2[1] : w2[0], b0 ? w2[0] : w2[1], 1);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = bfly_max_pair(b4 ? v[i + 8] : v[i], b4 ? v[i] : v[i + 8], 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) w4[i] = bfly_max_pair(b3 ? w8[i + 4] : w8[i], b3 ? w8[i] : w8[i + 4], 8);
#pragma unroll
        for (int i = 0; i < 2; ++i) w2[i] = bfly_max_pair(b2 ? w4[i + 2] : w4[i], b2 ? w4[i] : w4[i + 2], 4);
        r = bfly_max_pair(b1 ? w2[1] : w2[0], b1 ? w2[0] : w2[1], 2);
        r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, 1));
    }
    return r;
}

// Results go straight to global memory: a group that lives in one warp (pool 16 / 32) is stored,
// a group spread over 2 or 4 warps (pool 64 / 128) is combined with atomicMax on the float bits
// (all values >= 0 after ReLU; the caller zero-fills y first).
// The chunks are taken two at a time: both TMEM loads are issued before the single wait and the two
// shuffle butterflies are independent instruction streams the scheduler interleaves -- with two
// epilogue warps per scheduler the one-chunk version was bound by the latency of its 5-deep
// shuffle chain (the pooling epilogue, not the tensor pipe, set the tile period of the fused SA kernel).
__device__ __forceinline__ void pool_tile(uint32_t taddr, int ncols, int half, const float *bias_s, bool live, int pool,
                                          int lane, int q, long long tile, long long rows, int cout, int col_base,
                                          float *y, int ldy) {
    const bool all_live = __all_sync(0xffffffffu, live);
    // pooled row this lane writes, its validity and the lane's column inside a 16-column chunk (no
    // divisions: pool is 16, 32, 64 or 128)
    long long orow;
    int c;
    bool writer;
    const bool pool16 = pool == 16;
    if (pool16) { orow = tile * 8 + q * 2 + (lane >> 4); c = lane & 15; writer = true; }
    else {
        orow = pool == 32 ? tile * 4 + q : (pool == 64 ? tile * 2 + (q >> 1) : tile);
        c = lane >> 1;
        writer = (lane & 1) == 0;
    }
    writer = writer && orow * pool < rows;
    float *yrow = y + orow * ldy + col_base + c;
    auto emit = [&](float r, int c0) {
        const float o = fmaxf(r + bias_s[c0 + c], 0.f);
        if (writer && col_base + c0 + c < cout) {
            if (pool <= 32) yrow[c0] = o;
            else atomicMax(reinterpret_cast<unsigned int *>(yrow + c0), __float_as_uint(o));
        }
    };
    int c0 = half * 16;
    for (; c0 + 32 < ncols; c0 += 64) {
        uint32_t ra[16], rb[16];
        tmem_ld16(taddr + c0, ra);
        tmem_ld16(taddr + c0 + 32, rb);
        tmem_ld_wait();
        const float r0 = chunk_max(ra, live, all_live, pool16, lane);
        const float r1 = chunk_max(rb, live, all_live, pool16, lane);
        emit(r0, c0);
        emit(r1, c0 + 32);
    }
    if (c0 < ncols) {
        uint32_t ra[16];
        tmem_ld16(taddr + c0, ra);
        tmem_ld_wait();
        emit(chunk_max(ra, live, all_live, pool16, lane), c0);
    }
}

}  // namespace tc
