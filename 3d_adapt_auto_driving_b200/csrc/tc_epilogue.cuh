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

// max over the rows of each pooling group of act(acc + bias), for the chunks of this warp.
// Values are >= 0 after ReLU, so the float order is the unsigned order of the bits
// (redux.sync has no float form on sm_100).  part: [8][256] floats; for pool 16 the row index is
// 2*q + half-warp, otherwise q (pool 64 / 128 are combined across quadrants by pool_combine).
__device__ __forceinline__ void pool_tile(uint32_t taddr, int ncols, int half, const float *bias_s, bool live, int pool,
                                          int lane, int q, float *part) {
    for (int c0 = half * 16; c0 < ncols; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float o = fmaxf(__uint_as_float(v[j]) + bias_s[c0 + j], 0.f);
            const uint32_t u = live ? __float_as_uint(o) : 0u;
            uint32_t mx;
            if (pool == 16) {
                const uint32_t lo16 = __reduce_max_sync(0xffffffffu, lane < 16 ? u : 0u);
                const uint32_t hi16 = __reduce_max_sync(0xffffffffu, lane >= 16 ? u : 0u);
                mx = (lane & 16) ? hi16 : lo16;
            } else {
                mx = __reduce_max_sync(0xffffffffu, u);
            }
            if ((lane & 15) == j) mine = mx;
        }
        if (pool == 16) part[(q * 2 + (lane >> 4)) * 256 + c0 + (lane & 15)] = __uint_as_float(mine);
        else if (lane < 16) part[q * 256 + c0 + lane] = __uint_as_float(mine);
    }
}

// after a barrier among the epilogue warps: combine the quadrant maxima and write the pooled rows
__device__ __forceinline__ void pool_combine(const float *part, int pool, int ncols, long long tile, long long rows,
                                             int cout, int col_base, float *y, int ldy, int etid) {
    const int groups = kBM_ / pool;                       // per tile: 8, 4, 2 or 1
    const int parts_per_group = pool <= 32 ? 1 : pool / 32;
    for (int e = etid; e < groups * ncols; e += kEpiThreads) {
        const int g = e / ncols, c = e % ncols;
        const long long orow = tile * groups + g;
        if (orow * pool >= rows || col_base + c >= cout) continue;
        float mx = 0.f;
        if (pool == 16) mx = part[g * 256 + c];
        else
            for (int qq = 0; qq < parts_per_group; ++qq) mx = fmaxf(mx, part[(g * parts_per_group + qq) * 256 + c]);
        y[orow * ldy + col_base + c] = mx;
    }
}

}  // namespace tc
