// linear.cu -- fp32 shared-MLP layers (1x1 conv + folded BN + ReLU) with fused neighbourhood
// gather prologue and fused max-pool epilogue, for sm_100a.
//
// Replaces, for inference, the chain the reference builds out of separate library calls for
// every SA / FP / head layer (pointnet2_modules.py:37-48, pytorch_utils.py:5-101):
//   grouping_operation (xyz) -> subtract centre -> grouping_operation (features) -> cat ->
//   [cuDNN 1x1 conv -> BatchNorm -> ReLU] x L -> max_pool2d over nsample.
// None of the (B, C, npoint, nsample) intermediates of that chain are materialised here.
//
// Data layout: activations are POINT-major, X[row][channel] with a leading dimension, so a
// neighbour gather is a contiguous row read and layer outputs can be written straight into a
// column slice of a wider buffer (MSG scale concat and FP skip concat cost nothing).
//
// The first layer of an SA MLP acts on [xyz_j - centre ; feat_j].  It is split as
//   W1 [d ; f_j] + b = (W1f f_j + b)  +  W1x d        (exact algebra, different rounding order)
// so that H_j = W1f f_j + b is computed ONCE per source point (pn2_linear_f32) instead of once
// per (centre, neighbour) pair, and the pair-wise part is 3 FMAs per channel folded into the
// operand load of the second layer (GATHER prologue below).  For RCNN SA1 this removes a third
// of all FLOPs of the network.
//
// This file is the CUDA-core (FFMA) path: exact fp32 products, fp32 accumulation, used for every
// layer.  Arithmetic differs from cuDNN's only in summation order (tolerance 1e-4 rel in tests).
#include "common.cuh"

namespace {

constexpr int BM = 128;  // rows (points / pairs) per CTA tile
constexpr int BK = 16;
constexpr int kThreads = 256;

struct LinearParams {
    // operand A
    const float *x;        // [rows_x][ldx]  (GATHER: per-point H, rows addressed through idx)
    int ldx;
    int cin;               // true K
    long long rows;        // output rows before pooling (points, or centres*nsample)
    // GATHER prologue (all null/0 otherwise)
    const int32_t *idx;    // [rows] neighbour index inside its cloud
    const float *xyz;      // [clouds][n][3]
    const float *centres;  // [clouds][m][3]
    const float *wxyz;     // [3][cin]  rows: x, y, z coefficients of layer 1
    int n, m, ns;          // points per cloud, centres per cloud, samples per centre
    // operand B
    const float *w;        // [cout][ldw]
    int ldw;
    const float *bias;     // [cout] or null
    const float *res;      // [rows][ldr] added before the activation, or null (pool must be 1)
    int ldr;
    int cout;
    int relu;
    // output
    float *y;              // [rows / pool][ldy]
    int ldy;
    int pool;              // 1 = none, else max over `pool` consecutive rows (must divide BM, multiple of 4)
};

// TN = columns per thread; BN = 16*TN output channels per CTA tile.
template <int TN, bool GATHER>
__global__ void __launch_bounds__(kThreads) linear_kernel(const LinearParams p) {
    constexpr int BN = 16 * TN;
    constexpr int APAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + APAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;

    // ---- A loader: thread covers rows (tid>>2) and (tid>>2)+64, k offset (tid&3)*4 ----
    const int a_r = tid >> 2, a_k = (tid & 3) * 4;
    const float *a_ptr[2];
    float a_dx[2], a_dy[2], a_dz[2];
    bool a_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long r = row0 + a_r + h * 64;
        a_ok[h] = r < p.rows;
        a_ptr[h] = p.x;
        a_dx[h] = a_dy[h] = a_dz[h] = 0.f;
        if (a_ok[h]) {
            if (GATHER) {
                const long long centre = r / p.ns;          // global centre id = cloud*m + c
                const long long cloud = centre / p.m;
                const int j = __ldg(p.idx + r);
                const long long src = cloud * p.n + j;
                a_ptr[h] = p.x + src * p.ldx;
                const float *pj = p.xyz + src * 3;
                const float *pc = p.centres + centre * 3;
                a_dx[h] = __ldg(pj) - __ldg(pc);
                a_dy[h] = __ldg(pj + 1) - __ldg(pc + 1);
                a_dz[h] = __ldg(pj + 2) - __ldg(pc + 2);
            } else {
                a_ptr[h] = p.x + r * p.ldx;
            }
        }
    }
    // ---- B loader: BN rows of W, 4 threads per row like A; rows (tid>>2) [+64 when BN==128] ----
    constexpr int B_ITERS = (BN * BK / 4 + kThreads - 1) / kThreads;  // float4 loads per thread
    const bool vec_a = (p.ldx & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const bool vec_w = (p.ldw & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);

    float4 a_reg[2];
    float4 b_reg[B_ITERS];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + a_k;
            if (a_ok[h] && k < p.cin) {
                if (vec_a && k + 3 < p.cin) {
                    v = __ldg(reinterpret_cast<const float4 *>(a_ptr[h] + k));
                } else {
                    v.x = __ldg(a_ptr[h] + k);
                    if (k + 1 < p.cin) v.y = __ldg(a_ptr[h] + k + 1);
                    if (k + 2 < p.cin) v.z = __ldg(a_ptr[h] + k + 2);
                    if (k + 3 < p.cin) v.w = __ldg(a_ptr[h] + k + 3);
                }
                if (GATHER) {
                    // layer-1 output of this (centre, neighbour) pair: relu(H_j + W1x d)
                    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (k + q < p.cin) {
                            float t = o[q];
                            t = fmaf(__ldg(p.wxyz + k + q), a_dx[h], t);
                            t = fmaf(__ldg(p.wxyz + p.cin + k + q), a_dy[h], t);
                            t = fmaf(__ldg(p.wxyz + 2 * p.cin + k + q), a_dz[h], t);
                            o[q] = fmaxf(t, 0.f);
                        }
                    }
                    v = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            a_reg[h] = v;
        }
#pragma unroll
        for (int it = 0; it < B_ITERS; ++it) {
            const int e = tid + it * kThreads;  // float4 slot: row = e>>2, k4 = (e&3)*4
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int nrow = e >> 2, k = k0 + (e & 3) * 4;
            if (nrow < BN && col0 + nrow < p.cout && k < p.cin) {
                const float *wp = p.w + (long long)(col0 + nrow) * p.ldw + k;
                if (vec_w && k + 3 < p.cin) {
                    v = __ldg(reinterpret_cast<const float4 *>(wp));
                } else {
                    v.x = __ldg(wp);
                    if (k + 1 < p.cin) v.y = __ldg(wp + 1);
                    if (k + 2 < p.cin) v.z = __ldg(wp + 2);
                    if (k + 3 < p.cin) v.w = __ldg(wp + 3);
                }
            }
            b_reg[it] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = a_r + h * 64;
            As[buf][a_k + 0][r] = a_reg[h].x;
            As[buf][a_k + 1][r] = a_reg[h].y;
            As[buf][a_k + 2][r] = a_reg[h].z;
            As[buf][a_k + 3][r] = a_reg[h].w;
        }
#pragma unroll
        for (int it = 0; it < B_ITERS; ++it) {
            const int e = tid + it * kThreads;
            const int nrow = e >> 2, kk = (e & 3) * 4;
            if (nrow < BN) {
                Bs[buf][kk + 0][nrow] = b_reg[it].x;
                Bs[buf][kk + 1][nrow] = b_reg[it].y;
                Bs[buf][kk + 2][nrow] = b_reg[it].z;
                Bs[buf][kk + 3][nrow] = b_reg[it].w;
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int ktiles = (p.cin + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else {
                const float2 b0 = *reinterpret_cast<const float2 *>(&Bs[buf][kk][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, ReLU, optional max over `pool` consecutive rows ----
    int cols[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        if (TN == 8) cols[j] = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        else cols[j] = col0 + tx * TN + j;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const float bj = (p.bias && cols[j] < p.cout) ? __ldg(p.bias + cols[j]) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[i][j] + bj;
            if (p.relu && !p.res) v = fmaxf(v, 0.f);
            acc[i][j] = v;
        }
    }
    if (p.pool <= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (r >= p.rows) continue;
            float *yr = p.y + r * p.ldy;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (cols[j] < p.cout) {
                    float v = acc[i][j];
                    if (p.res) {
                        v += __ldg(p.res + r * p.ldr + cols[j]);
                        if (p.relu) v = fmaxf(v, 0.f);
                    }
                    yr[cols[j]] = v;
                }
        }
        return;
    }
    // pooled: thread-local max over each 4-row quad, then across quads through shared memory.
    // quad q (0..31) covers tile rows q*4..q*4+3 ; group g = q / (pool/4).
    float *red = &As[0][0][0];  // 32 quads x BN floats <= 16 KB, As is 2*16*132*4 = 16.9 KB
    __syncthreads();
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
        const int q = hq * 16 + ty;
        const long long rbase = row0 + q * 4;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (rbase + i < p.rows) mx = fmaxf(mx, acc[hq * 4 + i][j]);
            const int cl = (TN == 8) ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx * TN + j;
            red[q * BN + cl] = mx;
        }
    }
    __syncthreads();
    const int qpg = p.pool / 4;          // quads per group
    const int groups = BM / p.pool;      // groups per tile
    for (int e = tid; e < groups * BN; e += kThreads) {
        const int g = e / BN, cl = e % BN;
        const long long orow = row0 / p.pool + g;
        if (orow * p.pool >= p.rows || col0 + cl >= p.cout) continue;
        float mx = -INFINITY;
        for (int q = 0; q < qpg; ++q) mx = fmaxf(mx, red[(g * qpg + q) * BN + cl]);
        p.y[orow * p.ldy + col0 + cl] = mx;
    }
}

template <bool GATHER>
int launch_linear(const LinearParams &p, cudaStream_t stream) {
    if (p.rows <= 0 || p.cout <= 0) return PN2_OK;
    if (p.pool > 1 && (BM % p.pool != 0 || p.pool % 4 != 0)) {
        pn2_set_last_error("linear: pool must divide 128 and be a multiple of 4");
        return PN2_ERR_UNSUPPORTED;
    }
    const long long tiles = (p.rows + BM - 1) / BM;
    if (tiles > 2147483647LL) {
        pn2_set_last_error("linear: too many rows");
        return PN2_ERR_UNSUPPORTED;
    }
    if (p.cout > 64) {
        dim3 grid((unsigned)tiles, pn2_divup(p.cout, 128));
        linear_kernel<8, GATHER><<<grid, kThreads, 0, stream>>>(p);
    } else if (p.cout > 32) {
        dim3 grid((unsigned)tiles, 1);
        linear_kernel<4, GATHER><<<grid, kThreads, 0, stream>>>(p);
    } else {
        dim3 grid((unsigned)tiles, 1);
        linear_kernel<2, GATHER><<<grid, kThreads, 0, stream>>>(p);
    }
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

}  // namespace

// Y[r, 0:cout] = act( X[r, 0:cin] . W[c, 0:cin]^T + bias[c] [+ R[r, c]] ), optionally max-pooled over
// `pool` consecutive rows (pool = nsample for the last layer of an SA MLP).
// X (rows, ldx) ; W (cout, ldw) ; Y (rows/pool, ldy) -- all row-major f32 device buffers.
// Replaces pytorch_utils.py:23-31 (SharedMLP layer) [+ F.max_pool2d, pointnet2_modules.py:42].
PN2_API int pn2_linear_f32(const float *x, int ldx, const float *w, int ldw, const float *bias, const float *res,
                           int ldr, float *y, int ldy, long long rows, int cin, int cout, int relu, int pool,
                           cudaStream_t stream) {
    if (!x || !w || !y || rows < 0 || cin <= 0 || cout < 0 || ldx < cin || ldw < cin || pool < 1 ||
        (res && pool != 1)) {
        pn2_set_last_error("pn2_linear_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    LinearParams p = {};
    p.x = x; p.ldx = ldx; p.cin = cin; p.rows = rows;
    p.w = w; p.ldw = ldw; p.bias = bias; p.cout = cout; p.relu = relu;
    p.res = res; p.ldr = ldr;
    p.y = y; p.ldy = ldy; p.pool = pool;
    return launch_linear<false>(p, stream);
}

// Second layer of an SA MLP with the neighbourhood gather and the (split) first layer fused
// into the operand load:  for pair r = (centre c, sample s), j = idx[r]:
//   a_r[k] = relu( H[cloud, j, k] + W1x[:,k] . (xyz[cloud, j] - centres[cloud, c]) )
//   Y[r]   = act( a_r . W2^T + b2 ),  optionally max-pooled over the ns samples of a centre.
// h (clouds*n, ldh) per-point first-layer pre-activations (bias/BN folded in) ;
// idx (clouds, m, ns) int32 from ball_query ; wxyz (3, c1).
// Replaces QueryAndGroup.forward (pointnet2_utils.py:241-264) + two SharedMLP layers.
PN2_API int pn2_sa_group_linear_f32(const float *h, int ldh, const int32_t *idx, const float *xyz,
                                    const float *centres, const float *wxyz, const float *w, int ldw,
                                    const float *bias, float *y, int ldy, int clouds, int n, int m, int ns, int c1,
                                    int cout, int relu, int pool, cudaStream_t stream) {
    if (!h || !idx || !xyz || !centres || !wxyz || !w || !y || clouds < 0 || n <= 0 || m < 0 || ns <= 0 || c1 <= 0 ||
        ldh < c1 || ldw < c1 || pool < 1) {
        pn2_set_last_error("pn2_sa_group_linear_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    LinearParams p = {};
    p.x = h; p.ldx = ldh; p.cin = c1; p.rows = (long long)clouds * m * ns;
    p.idx = idx; p.xyz = xyz; p.centres = centres; p.wxyz = wxyz; p.n = n; p.m = m; p.ns = ns;
    p.w = w; p.ldw = ldw; p.bias = bias; p.cout = cout; p.relu = relu;
    p.y = y; p.ldy = ldy; p.pool = pool;
    return launch_linear<true>(p, stream);
}
