// common.cuh -- shared helpers for the sm_100a kernels of the PointRCNN hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PN2_API extern "C" __attribute__((visibility("default")))

// C-ABI status codes (include/pn2_b200.h)
#define PN2_OK 0
#define PN2_ERR_INVALID 1
#define PN2_ERR_LAUNCH 2
#define PN2_ERR_UNSUPPORTED 3

#define PN2_CHECK_LAUNCH()                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) {                            \
            pn2_set_last_error(cudaGetErrorString(e__));     \
            return PN2_ERR_LAUNCH;                           \
        }                                                    \
    } while (0)

void pn2_set_last_error(const char *msg);

static inline int pn2_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// Squared distance in the contraction order nvcc emits for the reference expression
// (dx*dx + dy*dy + dz*dz): t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t).
// Explicit _rn intrinsics so no compiler version can re-associate it.
__device__ __forceinline__ float pn2_sqdist(float dx, float dy, float dz) {
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    t = __fmaf_rn(dz, dz, t);
    return t;
}

__device__ __forceinline__ uint32_t pn2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
