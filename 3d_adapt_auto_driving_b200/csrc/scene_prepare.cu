// scene_prepare.cu -- the inference data path of KittiRCNNDataset.get_rpn_sample on the GPU, sm_100a.
//
// Replaces, per scene, the numpy chain of pointrcnn/lib/datasets/kitti_rcnn_dataset.py:249-320
//   calib.lidar_to_rect (lib/utils/calibration.py:51-58)  ->  calib.rect_to_img (:60-71)  ->
//   get_valid_flag (kitti_rcnn_dataset.py:201-222)  ->  pts_rect[valid]  ->  near / far index lists (:291-296)
// and the final pts_rect[choice] gather (:322), behind pn2_scene_filter_f32 / pn2_scene_gather_f32.
// It is row N1 of SURVEY.md 8(f): once the forward runs at >1700 scenes/s the CPU data path (7 ms per
// 60 k-point scene in one DataLoader worker) is the end-to-end bottleneck of eval_rcnn.py.
//
// What stays on the host, by construction: the np.random draws.  Which points a scene keeps is defined by
// the reference's MT19937 stream (np.random.choice / shuffle in a fixed order); those calls depend on the
// data only through three COUNTS (valid, near, far), so the host draws index-of-index arrays from the
// counts this kernel returns and the gather kernel resolves them (datasets/gpu_loader.py).
//
// Arithmetic is pinned to what numpy does on float32 inputs (checked bit for bit against the numpy path,
// tests/test_gpu_loader_gpu.py): np.dot of an (N,4) by (4,3) float32 matrix is, per output element,
//   fma(a3, b3, fma(a2, b2, fma(a1, b1, a0 * b0)))          (sgemm accumulates along k with FMAs)
// the image coordinates are IEEE float32 divisions by z (1e-9f where z == 0), the depth a float32
// subtraction, the image-bounds compares are float32 and the PC_AREA_SCOPE compares float64 (numpy
// promotes the float32 column against the float64 cfg scalars).
//
// One CTA per scene walks the raw cloud in index order, 1024 points per step; ballots + a shared-memory scan
// over the 32 warps give every valid point its position in the ORDERED compacted lists, exactly the order
// boolean-mask indexing and np.where produce.  No scratch, no atomics, one launch per batch of scenes.
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;

struct SceneCalib {       // per scene, 32 floats
    float m[12];          // lidar -> rect: np.dot(V2C.T, R0.T), (4,3) row-major
    float p[12];          // rect -> image: P2.T, (4,3) row-major
    float width, height;  // image size as float32 (numpy compares the float32 column with the python int)
    float pad[6];
};

__device__ __forceinline__ float dot4(float a0, float a1, float a2, const float *b, int j) {
    float t = __fmul_rn(a0, b[j]);
    t = __fmaf_rn(a1, b[3 + j], t);
    t = __fmaf_rn(a2, b[6 + j], t);
    return __fmaf_rn(1.0f, b[9 + j], t);
}

// raw: concatenated (x, y, z, intensity) float4 of all scenes; offsets[b] .. offsets[b+1] = scene b.
// valid (B, cap) float4 (rect x, y, z, intensity) in input order; near_list / far_list (B, cap) int32 = positions
// in `valid` of the points with z < 40 / z >= 40, ascending; counts (B, 4) int32 = {n_valid, n_near, n_far, 0}.
__global__ void __launch_bounds__(kThreads) scene_filter_kernel(const float4 *__restrict__ raw,
                                                               const long long *__restrict__ offsets,
                                                               const SceneCalib *__restrict__ calib,
                                                               double x0, double x1, double y0, double y1, double z0,
                                                               double z1, int reduce_by_range, float near_z,
                                                               float4 *__restrict__ valid, int32_t *__restrict__ near_list,
                                                               int32_t *__restrict__ far_list, int32_t *__restrict__ counts,
                                                               long long cap) {
    __shared__ int wv[kWarps], wn[kWarps];
    __shared__ int base_v, base_n;
    __shared__ SceneCalib c;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < (int)(sizeof(SceneCalib) / 4)) reinterpret_cast<float *>(&c)[tid] = reinterpret_cast<const float *>(calib + b)[tid];
    if (tid == 0) { base_v = 0; base_n = 0; }
    __syncthreads();
    const long long beg = offsets[b], n = offsets[b + 1] - beg;
    raw += beg;
    valid += (size_t)b * cap;
    near_list += (size_t)b * cap;
    far_list += (size_t)b * cap;
    const unsigned lt = (1u << lane) - 1u;

    for (long long i0 = 0; i0 < n; i0 += kThreads) {
        const long long i = i0 + tid;
        bool ok = false, is_near = false;
        float rx = 0.f, ry = 0.f, rz = 0.f, inten = 0.f;
        if (i < n) {
            const float4 q = __ldg(raw + i);
            inten = q.w;
            rx = dot4(q.x, q.y, q.z, c.m, 0);
            ry = dot4(q.x, q.y, q.z, c.m, 1);
            rz = dot4(q.x, q.y, q.z, c.m, 2);
            const float hx = dot4(rx, ry, rz, c.p, 0), hy = dot4(rx, ry, rz, c.p, 1), hz = dot4(rx, ry, rz, c.p, 2);
            const float zd = rz == 0.f ? 1e-9f : rz;
            const float u = __fdiv_rn(hx, zd), v = __fdiv_rn(hy, zd);
            const float depth = __fsub_rn(hz, c.p[11]);
            ok = u >= 0.f && u < c.width && v >= 0.f && v < c.height && depth >= 0.f;
            if (reduce_by_range)
                ok = ok && (double)rx >= x0 && (double)rx <= x1 && (double)ry >= y0 && (double)ry <= y1 &&
                     (double)rz >= z0 && (double)rz <= z1;
            is_near = rz < near_z;
        }
        const unsigned bv = __ballot_sync(0xffffffffu, ok);
        const unsigned bn = __ballot_sync(0xffffffffu, ok && is_near);
        if (lane == 0) { wv[warp] = __popc(bv); wn[warp] = __popc(bn); }
        __syncthreads();
        int pv = base_v, pn = base_n;            // exclusive prefix over the lower warps (32 warps: a short serial sum)
        for (int w = 0; w < warp; ++w) { pv += wv[w]; pn += wn[w]; }
        if (ok) {
            const int pos_v = pv + __popc(bv & lt);
            const int pos_n = pn + __popc(bn & lt);             // near points before this one
            valid[pos_v] = make_float4(rx, ry, rz, inten);
            if (is_near) near_list[pos_n] = pos_v;
            else far_list[pos_v - pos_n] = pos_v;               // far points before this one = valid - near
        }
        __syncthreads();
        if (tid == kThreads - 1) {                              // last warp's last lane holds the totals of this step
            base_v = pv + __popc(bv);
            base_n = pn + __popc(bn);
        }
        __syncthreads();
    }
    if (tid == 0) {
        counts[b * 4 + 0] = base_v;
        counts[b * 4 + 1] = base_n;
        counts[b * 4 + 2] = base_v - base_n;
        counts[b * 4 + 3] = 0;
    }
}

// sel (B, npoints) int32, the host's random choice after the shuffle, encoded against the lists above:
//   0 <= s < 2^30            : near_list[s]          (s-th near point)
//   2^30 <= s < 2^31         : far_list[s - 2^30]    (s-th far point)
//   s < 0                    : valid[-s - 1]         (direct position, the "fewer valid points than npoints" branch)
// pts (B, npoints, 3) = rect xyz of the chosen points; feat (B, npoints) = intensity - 0.5 (kitti_rcnn_dataset.py:323) or NULL.
__global__ void __launch_bounds__(256) scene_gather_kernel(const float4 *__restrict__ valid,
                                                          const int32_t *__restrict__ near_list,
                                                          const int32_t *__restrict__ far_list,
                                                          const int32_t *__restrict__ sel, float *__restrict__ pts,
                                                          float *__restrict__ feat, int npoints, long long cap) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npoints) return;
    const int s = __ldg(sel + (size_t)b * npoints + k);
    int pos;
    if (s < 0) pos = -s - 1;
    else if (s < (1 << 30)) pos = __ldg(near_list + (size_t)b * cap + s);
    else pos = __ldg(far_list + (size_t)b * cap + (s - (1 << 30)));
    const float4 q = __ldg(valid + (size_t)b * cap + pos);
    float *o = pts + ((size_t)b * npoints + k) * 3;
    o[0] = q.x; o[1] = q.y; o[2] = q.z;
    if (feat) feat[(size_t)b * npoints + k] = __fsub_rn(q.w, 0.5f);
}

}  // namespace

// calib: B x 32 floats (SceneCalib); scope: 6 doubles {x0, x1, y0, y1, z0, z1} by value.
PN2_API int pn2_scene_filter_f32(const float *raw, const long long *offsets, const float *calib, double x0, double x1,
                                 double y0, double y1, double z0, double z1, int reduce_by_range, float near_z,
                                 float *valid, int32_t *near_list, int32_t *far_list, int32_t *counts, int b,
                                 long long cap, cudaStream_t stream) {
    if (b < 0 || cap < 0 || (b > 0 && (!raw || !offsets || !calib || !valid || !near_list || !far_list || !counts)) ||
        (reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(valid) & 15)) {
        pn2_set_last_error("pn2_scene_filter_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0) return PN2_OK;
    scene_filter_kernel<<<b, kThreads, 0, stream>>>(reinterpret_cast<const float4 *>(raw), offsets,
                                                    reinterpret_cast<const SceneCalib *>(calib), x0, x1, y0, y1, z0, z1,
                                                    reduce_by_range, near_z, reinterpret_cast<float4 *>(valid), near_list,
                                                    far_list, counts, cap);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_scene_gather_f32(const float *valid, const int32_t *near_list, const int32_t *far_list,
                                 const int32_t *sel, float *pts, float *feat, int b, int npoints, long long cap,
                                 cudaStream_t stream) {
    if (b < 0 || npoints < 0 || (b * npoints > 0 && (!valid || !near_list || !far_list || !sel || !pts))) {
        pn2_set_last_error("pn2_scene_gather_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (b == 0 || npoints == 0) return PN2_OK;
    dim3 grid((npoints + 255) / 256, b);
    scene_gather_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(valid), near_list, far_list, sel, pts,
                                                  feat, npoints, cap);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same per-point pipeline as a HOST function, for the DataLoader workers of the unmodified eval_rcnn.py: the script
// owns the loop and the worker processes (tools/eval_rcnn.py:851-866), so the data path stays on the CPU there -- but it
// need not be ten numpy passes over a 120 000-point sweep (10-13 ms per scene, the bound of the whole script at 204
// scenes/s: profiles/r2_config5_n1.json).  One pass, the arithmetic of scene_filter_kernel (= numpy's, bit for bit):
//   rect = [x y z 1] . M,  image = [rect 1] . P,  u = hx / z, v = hy / z (z == 0 -> 1e-9f),  depth = hz - P[3][2],
//   valid = in the image, depth >= 0, (optionally) inside PC_AREA_SCOPE in float64.
// Outputs the compacted valid points (rect x, y, z, intensity) in input order and the positions of the near (z < near_z)
// and far ones among them, the three counts the np.random draws depend on.  FMAs are hardware FMAs when the CPU has them
// (every x86-64 server of the last decade; __builtin_cpu_supports), libm's fmaf otherwise: identical results.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// the loop body; __builtin_fmaf becomes a vfmadd instruction inside the target("fma") wrapper and a libm call in the other
static inline __attribute__((always_inline))
long long scene_filter_host(const float *raw, long long n, const float *m, const float *p, float width, float height,
                            int reduce_by_range, const double *scope, float near_z, float *valid, int32_t *near_list,
                            int32_t *far_list, long long *n_near_out) {
    long long nv = 0, nn = 0;
    const float p11 = p[11];
    for (long long i = 0; i < n; ++i) {
        const float x = raw[4 * i], y = raw[4 * i + 1], z = raw[4 * i + 2];
        float r[3], h[3];
        for (int j = 0; j < 3; ++j) {
            float t = x * m[j];
            t = __builtin_fmaf(y, m[3 + j], t);
            t = __builtin_fmaf(z, m[6 + j], t);
            r[j] = __builtin_fmaf(1.0f, m[9 + j], t);
        }
        for (int j = 0; j < 3; ++j) {
            float t = r[0] * p[j];
            t = __builtin_fmaf(r[1], p[3 + j], t);
            t = __builtin_fmaf(r[2], p[6 + j], t);
            h[j] = __builtin_fmaf(1.0f, p[9 + j], t);
        }
        const float zd = r[2] == 0.f ? 1e-9f : r[2];
        const float u = h[0] / zd, v = h[1] / zd;
        const float depth = h[2] - p11;
        bool ok = u >= 0.f && u < width && v >= 0.f && v < height && depth >= 0.f;
        if (ok && reduce_by_range)
            ok = (double)r[0] >= scope[0] && (double)r[0] <= scope[1] && (double)r[1] >= scope[2] && (double)r[1] <= scope[3] &&
                 (double)r[2] >= scope[4] && (double)r[2] <= scope[5];
        if (!ok) continue;
        valid[4 * nv] = r[0]; valid[4 * nv + 1] = r[1]; valid[4 * nv + 2] = r[2]; valid[4 * nv + 3] = raw[4 * i + 3];
        if (r[2] < near_z) near_list[nn++] = (int32_t)nv;
        else far_list[nv - nn] = (int32_t)nv;
        ++nv;
    }
    *n_near_out = nn;
    return nv;
}

__attribute__((target("fma"))) long long scene_filter_host_fma(const float *raw, long long n, const float *m, const float *p,
                                                              float width, float height, int reduce_by_range, const double *scope,
                                                              float near_z, float *valid, int32_t *near_list, int32_t *far_list,
                                                              long long *n_near_out) {
    return scene_filter_host(raw, n, m, p, width, height, reduce_by_range, scope, near_z, valid, near_list, far_list, n_near_out);
}
long long scene_filter_host_libm(const float *raw, long long n, const float *m, const float *p, float width, float height,
                                 int reduce_by_range, const double *scope, float near_z, float *valid, int32_t *near_list,
                                 int32_t *far_list, long long *n_near_out) {
    return scene_filter_host(raw, n, m, p, width, height, reduce_by_range, scope, near_z, valid, near_list, far_list, n_near_out);
}

}  // namespace

// HOST function (no device work, no stream): raw (n, 4) f32 sweep, m / p the (4, 3) row-major float32 matrices of
// calibration.Calibration (_velo_to_rect, P2.T), scope 6 doubles {x0, x1, y0, y1, z0, z1} -> valid (n, 4) f32 capacity,
// near_list / far_list (n) int32 capacity, counts[3] = {n_valid, n_near, n_far}.
PN2_API int pn2_scene_filter_host_f32(const float *raw, long long n, const float *m, const float *p, float width, float height,
                                      int reduce_by_range, const double *scope, float near_z, float *valid,
                                      int32_t *near_list, int32_t *far_list, long long *counts) {
    if (n < 0 || (n > 0 && (!raw || !valid || !near_list || !far_list)) || !m || !p || !counts || (reduce_by_range && !scope)) {
        pn2_set_last_error("pn2_scene_filter_host_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    long long nn = 0, nv;
    // contraction of `x * m[j]` + ... into FMAs would change the first product's rounding: the expressions above are
    // written so that only the explicit fma calls fuse (a lone product feeding an fma is not contractible further)
    if (__builtin_cpu_supports("fma")) nv = scene_filter_host_fma(raw, n, m, p, width, height, reduce_by_range, scope, near_z, valid, near_list, far_list, &nn);
    else nv = scene_filter_host_libm(raw, n, m, p, width, height, reduce_by_range, scope, near_z, valid, near_list, far_list, &nn);
    counts[0] = nv; counts[1] = nn; counts[2] = nv - nn;
    return PN2_OK;
}
