// spatial_order.cuh -- space-filling-curve ordering of the points of a cloud (scratch for the culled scans).
//
// ball_query.cu and interpolate.cu give the 32 query points of a warp a small common bounding
// box by walking the queries in Hilbert-curve order of their ground-plane (x, z) cell.  One CTA per
// cloud: bounding box, 64 x 64 cell histogram in shared memory, exclusive scan, scatter.  The
// order inside a cell is whatever the atomics produce -- any permutation is a valid input of
// the scans, whose results do not depend on it.
#pragma once
#include "common.cuh"

namespace {

constexpr int kOrderThreads = 1024;
constexpr int kOrderBits = 6;                          // per axis
constexpr int kOrderCells = 1 << (2 * kOrderBits);     // 4096

// position of cell (x, y) along the Hilbert curve of a 64 x 64 grid.  Unlike the Z-order curve the
// Hilbert curve has no jumps: consecutive positions are adjacent cells, so a run of consecutive
// points never straddles two distant parts of the scene.
__device__ __forceinline__ int order_hilbert(uint32_t x, uint32_t y) {
    constexpr uint32_t n = 1u << kOrderBits;
    uint32_t d = 0;
#pragma unroll
    for (uint32_t s = n >> 1; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx) { x = n - 1 - x; y = n - 1 - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return (int)d;
}

__global__ void __launch_bounds__(kOrderThreads) spatial_order_kernel(const float *__restrict__ pts,
                                                                      int32_t *__restrict__ order, int m) {
    __shared__ int hist[kOrderCells];
    __shared__ float red[4][kOrderThreads / 32];
    __shared__ int wsum[kOrderThreads / 32];
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *q = pts + (size_t)cloud * m * 3;
    order += (size_t)cloud * m;

    float xmin = 3.0e38f, xmax = -3.0e38f, zmin = 3.0e38f, zmax = -3.0e38f;
    for (int i = tid; i < m; i += kOrderThreads) {
        const float x = __ldg(q + (size_t)i * 3), z = __ldg(q + (size_t)i * 3 + 2);
        xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
        zmin = fminf(zmin, z); zmax = fmaxf(zmax, z);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
        zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    }
    if (lane == 0) { red[0][warp] = xmin; red[1][warp] = xmax; red[2][warp] = zmin; red[3][warp] = zmax; }
    for (int i = tid; i < kOrderCells; i += kOrderThreads) hist[i] = 0;
    __syncthreads();
    for (int w = 0; w < kOrderThreads / 32; ++w) {
        xmin = fminf(xmin, red[0][w]); xmax = fmaxf(xmax, red[1][w]);
        zmin = fminf(zmin, red[2][w]); zmax = fmaxf(zmax, red[3][w]);
    }
    const float top = (float)((1 << kOrderBits) - 1);
    const float sx = xmax > xmin ? top / (xmax - xmin) : 0.f;
    const float sz = zmax > zmin ? top / (zmax - zmin) : 0.f;
    auto cell_of = [&](int i) -> int {
        const float x = __ldg(q + (size_t)i * 3), z = __ldg(q + (size_t)i * 3 + 2);
        // non-finite coordinates land in some cell; only the ORDER depends on it
        const float fx = fminf(fmaxf((x - xmin) * sx, 0.f), top), fz = fminf(fmaxf((z - zmin) * sz, 0.f), top);
        const uint32_t cx = (uint32_t)(int)fx & ((1u << kOrderBits) - 1u), cz = (uint32_t)(int)fz & ((1u << kOrderBits) - 1u);
        return order_hilbert(cx, cz);
    };
    for (int i = tid; i < m; i += kOrderThreads) atomicAdd(&hist[cell_of(i)], 1);
    __syncthreads();
    // exclusive scan of the 4096 counters: 4 per thread, warp scan, scan of the 32 warp sums
    constexpr int kPer = kOrderCells / kOrderThreads;
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
        int w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
    }
    __syncthreads();
    int excl = wsum[warp] + inc - run;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { hist[tid * kPer + j] = excl; excl += v[j]; }
    __syncthreads();
    for (int i = tid; i < m; i += kOrderThreads) order[atomicAdd(&hist[cell_of(i)], 1)] = i;
}

// pts (b, m, 3) -> order (b, m) int32: a permutation of 0..m-1 per cloud
inline cudaError_t launch_spatial_order(const float *pts, int32_t *order, int b, int m, cudaStream_t stream) {
    spatial_order_kernel<<<b, kOrderThreads, 0, stream>>>(pts, order, m);
    return cudaGetLastError();
}

}  // namespace
