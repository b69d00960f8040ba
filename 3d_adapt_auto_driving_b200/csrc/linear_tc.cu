// linear_tc.cu -- shared-MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as pn2_linear_f32 / pn2_sa_group_linear_f32 (linear.cu): one SharedMLP layer
//   Y = act(X . W^T + b [+ R]) [max over `pool` consecutive rows]
// (pytorch_utils.py:5-101 conv1x1 + folded BN + ReLU, pointnet2_modules.py:42 max_pool2d), optionally with
// the QueryAndGroup gather and the pair-wise half of the previous layer fused into the operand
// load (pointnet2_utils.py:241-264), but the contraction runs as tcgen05.mma.kind::f16 with the
// accumulator in tensor memory.
//
// fp32 semantics on bf16 tensor cores: every operand is split as x = hi + lo (two bf16, 16
// significant bits) and the product is accumulated in fp32 as hi.hi + hi.lo + lo.hi ("BF16x3").
// Measured error <= 7e-6 of the tensor scale per layer (tests/test_mlp_modules_gpu.py), inside the
// 1e-4 relative tolerance of BASELINE.json; plain TF32 (1e-3) is not.
//
// Structure (persistent, one CTA per SM, 13 warps):
//   warps 0-3   epilogue: tcgen05.ld accumulator -> bias / residual / ReLU -> coalesced store, or
//               max over nsample via redux.sync on the (non-negative) float bits;
//   warp  4     TMEM allocation; one elected lane issues tcgen05.mma (3 per 16-wide k step) and
//               tcgen05.commit onto the stage-empty / accumulator-full mbarriers;
//   warps 5-12  producers: global fp32 rows (plain or gathered + layer-1 xyz term + ReLU) ->
//               bf16 hi/lo -> 128B-swizzled K-major shared tiles, two K-blocks of loads in flight
//               per thread; the weight K-block (pre-split, pre-swizzled on the host) arrives by
//               one cp.async.bulk (TMA bulk copy) per stage on the same mbarrier.
// The accumulator is double-buffered in TMEM (2 x N <= 512 columns), so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int BM = 128;          // rows per tile = UMMA M
constexpr int BK = 64;           // bf16 per K-block row = one 128-byte swizzle atom
constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMaxStages = 4;
constexpr int kABytes = BM * BK * 2;  // one bf16 A tile (hi or lo) = 16 KB
constexpr int kStgLd = 17;            // epilogue staging row stride (floats)

struct TcParams {
    const float *x; int ldx; int cin; long long rows;
    const int32_t *idx; const float *xyz; const float *centres; const float *wxyz; int n, m, ns;
    const uint8_t *wblob;   // [nchunks][nkb][hi tile | lo tile], each ntile x 128 B, swizzled
    const float *bias; const float *res; int ldr; int cout; int relu;
    float *y; int ldy; int pool;
    int ntile, nchunks, nkb, stages;
    long long items;        // tiles * nchunks
    int vec_ok;             // rows of x are 16-byte aligned (ldx % 4 == 0, base aligned)
};

struct RowMeta {
    int src;                // source row of x, or -1 (row beyond the end: zeros)
    float dx, dy, dz;
};

// shared memory carve-up (dynamic, 1024-aligned base)
struct SmemLayout {
    uint32_t stage_bytes, off_meta, off_wx, off_bias, off_stg, off_part, off_bars, off_tmem, total;
};
__host__ __device__ inline SmemLayout make_layout(int ntile, int stages, int kpad, bool gather) {
    SmemLayout L;
    L.stage_bytes = 2 * kABytes + 2 * ntile * 128;
    uint32_t o = L.stage_bytes * stages;
    L.off_meta = o; o += 2 * BM * sizeof(RowMeta);
    L.off_wx = o;   o += gather ? 3 * kpad * 4 : 0;
    L.off_bias = o; o += 256 * 4;
    L.off_stg = o;  o += kEpiWarps * 32 * kStgLd * 4;
    L.off_part = o; o += 8 * 256 * 4;
    L.off_bars = o; o += (2 * kMaxStages + 4) * 8;
    L.off_tmem = o; o += 16;
    L.total = o;
    return L;
}

template <bool GATHER>
__global__ void __launch_bounds__(kThreads, 1) linear_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled operand tiles need 1 KB alignment in the shared window (1 KB of slack is allocated)
    uint8_t *smem = smem_raw + ((1024u - (pn2_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = p.nkb * BK;
    const SmemLayout L = make_layout(p.ntile, p.stages, kpad, GATHER);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint64_t *empty = full + kMaxStages;
    uint64_t *acc_full = empty + kMaxStages;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_tmem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], kProdWarps + 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
    if (GATHER) {
        float *wxs = reinterpret_cast<float *>(smem + L.off_wx);
        for (int i = threadIdx.x; i < 3 * kpad; i += kThreads) {
            const int c = i / kpad, k = i % kpad;
            wxs[i] = k < p.cin ? __ldg(p.wxyz + c * p.cin + k) : 0.f;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= kMmaWarp + 1) {
        // =============================== producers ===============================
        const int ptid = threadIdx.x - (kMmaWarp + 1) * 32;   // 0..255
        const int pw = ptid >> 5;
        const int rsub = pw * 2 + (lane >> 4);                 // row within a 16-row pass
        const int kq = (lane & 15) * 4;                        // first k of this thread's float4 in a K-block
        RowMeta *meta = reinterpret_cast<RowMeta *>(smem + L.off_meta);
        const float *wxs = reinterpret_cast<const float *>(smem + L.off_wx);

        auto fill_meta = [&](long long item, int buf) {
            if (ptid < BM) {
                RowMeta mt;
                mt.src = -1; mt.dx = mt.dy = mt.dz = 0.f;
                if (item < p.items) {
                    const long long r = (item / p.nchunks) * BM + ptid;
                    if (r < p.rows) {
                        if (GATHER) {
                            const long long centre = r / p.ns, cloud = centre / p.m;
                            const int j = __ldg(p.idx + r);
                            const long long src = cloud * p.n + j;
                            const float *pj = p.xyz + src * 3, *pc = p.centres + centre * 3;
                            mt.src = (int)src;
                            mt.dx = __ldg(pj) - __ldg(pc);
                            mt.dy = __ldg(pj + 1) - __ldg(pc + 1);
                            mt.dz = __ldg(pj + 2) - __ldg(pc + 2);
                        } else {
                            mt.src = (int)r;
                        }
                    }
                }
                meta[buf * BM + ptid] = mt;
            }
        };
        // flat sequence of steps t = (item iteration, kb); loads run two steps ahead of stores
        const long long first = blockIdx.x, stride = gridDim.x;
        const long long my_items = first < p.items ? (p.items - first + stride - 1) / stride : 0;
        const long long total_steps = my_items * p.nkb;

        float4 areg0[8], areg1[8];   // two K-blocks of loads in flight; kept in registers (static slot)
        auto issue_loads = [&](long long t, float4 (&areg)[8]) {
            const long long it = t / p.nkb;
            const int kb = (int)(t % p.nkb);
            const RowMeta *mt = meta + (it & 1) * BM;
            const int k = kb * BK + kq;
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
                const int src = mt[ps * 16 + rsub].src;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src >= 0 && k < p.cin) {
                    const float *px = p.x + (long long)src * p.ldx + k;
                    if (p.vec_ok && k + 3 < p.cin) {
                        v = __ldg(reinterpret_cast<const float4 *>(px));
                    } else {
                        v.x = __ldg(px);
                        if (k + 1 < p.cin) v.y = __ldg(px + 1);
                        if (k + 2 < p.cin) v.z = __ldg(px + 2);
                        if (k + 3 < p.cin) v.w = __ldg(px + 3);
                    }
                }
                areg[ps] = v;
            }
        };

        if (my_items > 0) {
            fill_meta(first, 0);
            fill_meta(first + stride, 1);
            named_bar_sync(1, kProdThreads);
            issue_loads(0, areg0);
            if (total_steps > 1) issue_loads(1, areg1);
        }
        int stage = 0;
        uint32_t phase = 0;
        auto do_step = [&](long long t, float4 (&areg)[8]) {
            const long long it = t / p.nkb;
            const int kb = (int)(t % p.nkb);
            const long long item = first + it * stride;
            uint8_t *sbase = smem + (size_t)stage * L.stage_bytes;
            mbar_wait(&empty[stage], phase ^ 1);
            if (ptid == 0) {
                // weight K-block: one bulk copy, completes on the same barrier as the A rows
                const uint32_t bytes = 2u * p.ntile * 128u;
                const uint8_t *src = p.wblob + ((size_t)(item % p.nchunks) * p.nkb + kb) * bytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pn2_smem_u32(&full[stage])),
                             "r"(bytes)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        pn2_smem_u32(sbase + 2 * kABytes)),
                    "l"(src), "r"(bytes), "r"(pn2_smem_u32(&full[stage]))
                    : "memory");
            }
            // transform + store this thread's 8 row segments
            const RowMeta *mt = meta + (it & 1) * BM;
            const int k = kb * BK + kq;
            float4 w0, w1, w2;
            if (GATHER) {
                w0 = *reinterpret_cast<const float4 *>(wxs + k);
                w1 = *reinterpret_cast<const float4 *>(wxs + kpad + k);
                w2 = *reinterpret_cast<const float4 *>(wxs + 2 * kpad + k);
            }
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
                const int r = ps * 16 + rsub;
                float4 v = areg[ps];
                if (GATHER) {
                    const RowMeta q = mt[r];
                    if (q.src >= 0) {
                        v.x = fmaxf(fmaf(w2.x, q.dz, fmaf(w1.x, q.dy, fmaf(w0.x, q.dx, v.x))), 0.f);
                        v.y = fmaxf(fmaf(w2.y, q.dz, fmaf(w1.y, q.dy, fmaf(w0.y, q.dx, v.y))), 0.f);
                        v.z = fmaxf(fmaf(w2.z, q.dz, fmaf(w1.z, q.dy, fmaf(w0.z, q.dx, v.z))), 0.f);
                        v.w = fmaxf(fmaf(w2.w, q.dz, fmaf(w1.w, q.dy, fmaf(w0.w, q.dx, v.w))), 0.f);
                        if (k + 3 >= p.cin) {   // zero the K padding again (bias-like terms must not leak)
                            if (k + 0 >= p.cin) v.x = 0.f;
                            if (k + 1 >= p.cin) v.y = 0.f;
                            if (k + 2 >= p.cin) v.z = 0.f;
                            v.w = 0.f;
                        }
                    }
                }
                uint2 hi, lo;
                split2(v.x, v.y, hi.x, lo.x);
                split2(v.z, v.w, hi.y, lo.y);
                const uint32_t off = sw128_offset(r, (lane & 15) >> 1) + ((lane & 1) << 3);
                *reinterpret_cast<uint2 *>(sbase + off) = hi;
                *reinterpret_cast<uint2 *>(sbase + kABytes + off) = lo;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
            // refill: loads of step t+2; at an item boundary publish the meta of the item after next
            if (kb == p.nkb - 1) {
                // every producer is done reading meta[it&1] once it passes this barrier
                named_bar_sync(1, kProdThreads);
                fill_meta(first + (it + 2) * stride, (int)(it & 1));
                named_bar_sync(1, kProdThreads);
            }
            if (t + 2 < total_steps) issue_loads(t + 2, areg);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        };
        for (long long t = 0; t < total_steps; t += 2) {
            do_step(t, areg0);
            if (t + 1 < total_steps) do_step(t + 1, areg1);
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, p.ntile);
            int stage = 0;
            uint32_t phase = 0;
            long long it = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int a = (int)(it & 1);
                mbar_wait(&acc_empty[a], (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.ntile);
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t sa = pn2_smem_u32(smem + (size_t)stage * L.stage_bytes);
                    const uint64_t a_hi = make_smem_desc_sw128(sa);
                    const uint64_t a_lo = make_smem_desc_sw128(sa + kABytes);
                    const uint64_t b_hi = make_smem_desc_sw128(sa + 2 * kABytes);
                    const uint64_t b_lo = make_smem_desc_sw128(sa + 2 * kABytes + p.ntile * 128);
                    const int krem = p.cin - kb * BK;
                    const int ksteps = krem >= BK ? 4 : (krem + 15) >> 4;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes >> 4 along K inside the atom
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        mma_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, acc);
                        mma_ss(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                        mma_ss(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                    }
                    mma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                mma_commit(&acc_full[a]);
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        float *bias_s = reinterpret_cast<float *>(smem + L.off_bias);
        float *stg = reinterpret_cast<float *>(smem + L.off_stg) + warp * 32 * kStgLd;
        float *part = reinterpret_cast<float *>(smem + L.off_part);
        const int etid = threadIdx.x;   // 0..127
        long long it = 0;
        int cur_chunk = -1;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int a = (int)(it & 1);
            const long long tile = item / p.nchunks;
            const int chunk = (int)(item % p.nchunks);
            const int col_base = chunk * p.ntile;
            if (chunk != cur_chunk) {
                named_bar_sync(2, kEpiWarps * 32);
                for (int c = etid; c < p.ntile; c += kEpiWarps * 32)
                    bias_s[c] = (p.bias && col_base + c < p.cout) ? __ldg(p.bias + col_base + c) : 0.f;
                named_bar_sync(2, kEpiWarps * 32);
                cur_chunk = chunk;
            }
            mbar_wait(&acc_full[a], (uint32_t)((it >> 1) & 1));
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * p.ntile);
            const long long row0 = tile * BM + warp * 32;
            for (int c0 = 0; c0 < p.ntile; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (p.pool <= 1) {
                    // transpose through shared memory: lane = row  ->  lanes = 2 rows x 16 columns
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[lane * kStgLd + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int cc = lane & 15;
                    const int col = col_base + c0 + cc;
                    const float bj = bias_s[c0 + cc];
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) {
                        const int rr = i * 2 + (lane >> 4);
                        const long long r = row0 + rr;
                        if (r < p.rows && col < p.cout) {
                            float o = stg[rr * kStgLd + cc] + bj;
                            if (p.res) o += __ldg(p.res + r * p.ldr + col);
                            if (p.relu) o = fmaxf(o, 0.f);
                            p.y[r * p.ldy + col] = o;
                        }
                    }
                    __syncwarp();
                } else {
                    // max over the rows of a group; values are >= 0 after ReLU so the float order is
                    // the unsigned order of the bits (redux.sync has no float form on sm_100)
                    const bool live = row0 + lane < p.rows;
                    uint32_t mine = 0u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float o = fmaxf(__uint_as_float(v[j]) + bias_s[c0 + j], 0.f);
                        const uint32_t u = live ? __float_as_uint(o) : 0u;
                        uint32_t mx;
                        if (p.pool == 16) {
                            const uint32_t lo16 = __reduce_max_sync(0xffffffffu, lane < 16 ? u : 0u);
                            const uint32_t hi16 = __reduce_max_sync(0xffffffffu, lane >= 16 ? u : 0u);
                            mx = (lane & 16) ? hi16 : lo16;
                        } else {
                            mx = __reduce_max_sync(0xffffffffu, u);
                        }
                        if ((lane & 15) == j) mine = mx;
                    }
                    // lane j (< 16) holds column c0+j of segment 0, lane 16+j of segment 1 (pool 16)
                    if (p.pool == 16) part[(warp * 2 + (lane >> 4)) * 256 + c0 + (lane & 15)] = __uint_as_float(mine);
                    else if (lane < 16) part[warp * 256 + c0 + lane] = __uint_as_float(mine);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            if (p.pool > 1) {
                named_bar_sync(2, kEpiWarps * 32);
                const int groups = BM / p.pool;                    // per tile: 8, 4, 2 or 1
                const int parts_per_group = p.pool <= 32 ? 1 : p.pool / 32;
                for (int e = etid; e < groups * p.ntile; e += kEpiWarps * 32) {
                    const int g = e / p.ntile, c = e % p.ntile;
                    const long long orow = tile * groups + g;
                    if (orow * p.pool >= p.rows || col_base + c >= p.cout) continue;
                    float mx = 0.f;
                    if (p.pool == 16) mx = part[g * 256 + c];
                    else
                        for (int q = 0; q < parts_per_group; ++q) mx = fmaxf(mx, part[(g * parts_per_group + q) * 256 + c]);
                    p.y[orow * p.ldy + col_base + c] = mx;
                }
                named_bar_sync(2, kEpiWarps * 32);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

int g_sm_count = 0;
int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <bool GATHER>
int launch_tc(TcParams &p, cudaStream_t stream) {
    const int kpad = p.nkb * BK;
    int stages = kMaxStages;
    SmemLayout L = make_layout(p.ntile, stages, kpad, GATHER);
    while (stages > 2 && L.total + 1024 > 227 * 1024) L = make_layout(p.ntile, --stages, kpad, GATHER);
    if (L.total + 1024 > 227 * 1024) {
        pn2_set_last_error("linear_tc: shared memory budget exceeded");
        return PN2_ERR_UNSUPPORTED;
    }
    p.stages = stages;
    static bool attr_done[2] = {false, false};
    if (!attr_done[GATHER]) {
        cudaFuncSetAttribute(linear_tc_kernel<GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_done[GATHER] = true;
    }
    const long long grid = p.items < sm_count() ? p.items : sm_count();
    linear_tc_kernel<GATHER><<<(unsigned)grid, kThreads, L.total + 1024, stream>>>(p);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

int fill_common(TcParams &p, const void *wblob, int ntile, int nchunks, int nkb, const float *bias, const float *res,
                int ldr, float *y, int ldy, long long rows, int cin, int cout, int relu, int pool) {
    if (!wblob || !y || rows < 0 || cin <= 0 || cout <= 0 || pool < 1 || ntile < 16 || ntile > 256 || (ntile & 15) ||
        nchunks < 1 || nkb < 1 || nkb * BK < cin || (long long)nchunks * ntile < cout || (res && pool != 1)) {
        pn2_set_last_error("linear_tc: bad argument");
        return PN2_ERR_INVALID;
    }
    if (pool > 1 && !(pool == 16 || pool == 32 || pool == 64 || pool == 128)) {
        pn2_set_last_error("linear_tc: pool must be 16, 32, 64 or 128");
        return PN2_ERR_UNSUPPORTED;
    }
    if (pool > 1 && !relu) {
        pn2_set_last_error("linear_tc: the fused max needs a ReLU layer (non-negative values)");
        return PN2_ERR_UNSUPPORTED;
    }
    p.wblob = static_cast<const uint8_t *>(wblob);
    p.ntile = ntile; p.nchunks = nchunks; p.nkb = nkb;
    p.bias = bias; p.res = res; p.ldr = ldr; p.cout = cout; p.relu = relu;
    p.y = y; p.ldy = ldy; p.pool = pool; p.cin = cin; p.rows = rows;
    p.items = ((rows + BM - 1) / BM) * nchunks;
    return PN2_OK;
}

}  // namespace

// Tensor-core version of pn2_linear_f32.  wblob: weights split into bf16 hi/lo and laid out by
// fused.pack_tc() as [nchunks][nkb][hi tile | lo tile], tiles of ntile x 64 bf16 in the 128B-swizzled
// K-major order the MMA reads.
PN2_API int pn2_linear_tc_f32(const float *x, int ldx, const void *wblob, int ntile, int nchunks, int nkb,
                              const float *bias, const float *res, int ldr, float *y, int ldy, long long rows, int cin,
                              int cout, int relu, int pool, cudaStream_t stream) {
    TcParams p = {};
    if (!x || ldx < cin) {
        pn2_set_last_error("pn2_linear_tc_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    const int rc = fill_common(p, wblob, ntile, nchunks, nkb, bias, res, ldr, y, ldy, rows, cin, cout, relu, pool);
    if (rc) return rc;
    if (rows == 0) return PN2_OK;
    if (rows > 2147483647LL) {
        pn2_set_last_error("pn2_linear_tc_f32: too many rows");
        return PN2_ERR_UNSUPPORTED;
    }
    p.x = x; p.ldx = ldx;
    p.vec_ok = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    return launch_tc<false>(p, stream);
}

// Tensor-core version of pn2_sa_group_linear_f32 (gather + xyz half of layer 1 + ReLU fused into
// the operand load of layer 2).
PN2_API int pn2_sa_group_linear_tc_f32(const float *h, int ldh, const int32_t *idx, const float *xyz,
                                       const float *centres, const float *wxyz, const void *wblob, int ntile,
                                       int nchunks, int nkb, const float *bias, float *y, int ldy, int clouds, int n,
                                       int m, int ns, int c1, int cout, int relu, int pool, cudaStream_t stream) {
    TcParams p = {};
    if (!h || !idx || !xyz || !centres || !wxyz || clouds < 0 || n <= 0 || m < 0 || ns <= 0 || ldh < c1 ||
        (long long)clouds * n > 2147483647LL) {
        pn2_set_last_error("pn2_sa_group_linear_tc_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    const long long rows = (long long)clouds * m * ns;
    const int rc = fill_common(p, wblob, ntile, nchunks, nkb, bias, nullptr, 0, y, ldy, rows, c1, cout, relu, pool);
    if (rc) return rc;
    if (rows == 0) return PN2_OK;
    p.x = h; p.ldx = ldh;
    p.idx = idx; p.xyz = xyz; p.centres = centres; p.wxyz = wxyz; p.n = n; p.m = m; p.ns = ns;
    p.vec_ok = ((ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(h) & 15) == 0);
    return launch_tc<true>(p, stream);
}
