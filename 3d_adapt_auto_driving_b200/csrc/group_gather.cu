// group_gather.cu -- index gathers (and their scatter-add backward) for sm_100a.
//
// Replaces sampling_gpu.cu:8-63 (gather_points[_grad]) and group_points_gpu.cu:8-86
// (group_points[_grad]) of pointrcnn/pointnet2_lib/pointnet2/src/.  Pure data movement: the
// index stream is read once, coalesced, and reused for a block of channels held in registers
// (the reference re-reads idx for every channel: grid.y = C); the output is written coalesced.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChan = 8;  // channels per thread pass

// out[b,c,e] = points[b,c,idx[b,e]]   e in [0,E)   (E = M for gather, M*ns for group)
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const float *__restrict__ points,
                                                              const int32_t *__restrict__ idx, float *__restrict__ out,
                                                              int c, int n, long long e_total) {
    const int cloud = blockIdx.z;
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= e_total) return;
    const int c0 = blockIdx.y * kChan;
    const int k = __ldg(idx + (size_t)cloud * e_total + e);
    const float *src = points + ((size_t)cloud * c + c0) * n + k;
    float *dst = out + ((size_t)cloud * c + c0) * e_total + e;
    float v[kChan];
#pragma unroll
    for (int j = 0; j < kChan; ++j)
        if (c0 + j < c) v[j] = __ldg(src + (size_t)j * n);
#pragma unroll
    for (int j = 0; j < kChan; ++j)
        if (c0 + j < c) dst[(size_t)j * e_total] = v[j];
}

// grad_points[b,c,idx[b,e]] += grad_out[b,c,e]
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(const float *__restrict__ grad_out,
                                                               const int32_t *__restrict__ idx,
                                                               float *__restrict__ grad_points, int c, int n,
                                                               long long e_total) {
    const int cloud = blockIdx.z;
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= e_total) return;
    const int c0 = blockIdx.y * kChan;
    const int k = __ldg(idx + (size_t)cloud * e_total + e);
    const float *src = grad_out + ((size_t)cloud * c + c0) * e_total + e;
    float *dst = grad_points + ((size_t)cloud * c + c0) * n + k;
#pragma unroll
    for (int j = 0; j < kChan; ++j)
        if (c0 + j < c) atomicAdd(dst + (size_t)j * n, __ldg(src + (size_t)j * e_total));
}

int launch_rows(bool scatter, const float *a, const int32_t *idx, float *o, int b, int c, int n, long long e,
                cudaStream_t s) {
    if (b < 0 || c < 0 || n < 0 || e < 0) {
        pn2_set_last_error("gather/group: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || c == 0 || e == 0) return PN2_OK;
    dim3 grid(pn2_divup(e, kThreads), pn2_divup(c, kChan), b);
    if (grid.y > 65535 || grid.z > 65535) {
        pn2_set_last_error("gather/group: batch or channel count too large");
        return PN2_ERR_UNSUPPORTED;
    }
    if (scatter) scatter_rows_kernel<<<grid, kThreads, 0, s>>>(a, idx, o, c, n, e);
    else gather_rows_kernel<<<grid, kThreads, 0, s>>>(a, idx, o, c, n, e);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

}  // namespace

PN2_API int pn2_gather_points_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m,
                                  cudaStream_t stream) {
    return launch_rows(false, points, idx, out, b, c, n, m, stream);
}
PN2_API int pn2_gather_points_grad_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c,
                                       int n, int m, cudaStream_t stream) {
    return launch_rows(true, grad_out, idx, grad_points, b, c, n, m, stream);
}
PN2_API int pn2_group_points_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m,
                                 int nsample, cudaStream_t stream) {
    return launch_rows(false, points, idx, out, b, c, n, (long long)m * nsample, stream);
}
PN2_API int pn2_group_points_grad_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c,
                                      int n, int m, int nsample, cudaStream_t stream) {
    return launch_rows(true, grad_out, idx, grad_points, b, c, n, (long long)m * nsample, stream);
}
