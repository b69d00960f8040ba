// sa_fused_t_tc.cu -- the fused set-abstraction MLP (sa_fused_tc.cu) with a TRANSPOSED last layer.
//
// Same contract as pn2_sa_fused_tc_f32: QueryAndGroup (pointnet2_utils.py:241-264) + SharedMLP layers
// 1-3 (pytorch_utils.py:5-101) + F.max_pool2d over nsample (pointnet2_modules.py:42) for one (SA module,
// scale), given the per-point half H of layer 1 -- for the shapes whose last layer has 128 output
// channels (RCNN SA1 [128,128,128], RPN SA2 [64,64,128] / [64,96,128]) or 256 (RCNN SA2 [128,128,256], which
// does not fit the row-major kernel at all and used to run layer by layer through a 1.6 GB intermediate).
//
// Why a second kernel.  In sa_fused_tc.cu the layer-3 accumulator has one ROW of the tile per TMEM lane,
// so the max over the nsample rows of a centre is a cross-lane reduction: a 5-level shuffle butterfly
// per 16-column chunk, ~85 issued instructions per chunk and a dependent SHFL chain that the two
// epilogue warps of a scheduler cannot hide.  The in-kernel stopwatch (tools/prof_tc.py) showed that
// pooling epilogue (3.4 k cycles per tile) plus the single-buffered layer-3 accumulator it drains
// serialised against the layer-3 MMAs: 6.6 k cycles per tile against 3.7 k of tensor-pipe work.
// Here layer 3 is computed as  acc3^T (channels x rows) = W3 (A operand, resident in TENSOR MEMORY)
// x A2^T (B operand, the layer-2 activation tile in shared memory):
//   * a TMEM lane now holds one output CHANNEL and the 128 columns are the tile's rows, so the max over
//     nsample consecutive rows is a plain in-thread FMNMX chain over the registers of a tcgen05.ld --
//     no shuffles, no atomics, no zero-fill of the output, and the pooled row is written as one coalesced
//     128-byte store per warp (lanes = consecutive channels);
//   * W3 leaves shared memory (it is loaded into TMEM once per CTA), which pays for the activation tile
//     A2 (bf16 hi/lo, 64 KB) that the layer-2 epilogue now writes to shared memory instead of TMEM;
//   * TMEM: acc2[0] | acc2[1] | W3.hi | W3.lo | acc3  =  2 n2 + c2 + 128 <= 512 columns: the LAYER-2
//     accumulator is double-buffered, so layer 2 of tile i+2 runs on the tensor pipe while the epilogue
//     converts tile i+1, and the (now ~10x cheaper) pooling epilogue never holds up layer 3.
//     With 256 output channels layer 3 runs as two passes of 128 over the same A2 tile (W3 takes 2 c2 columns,
//     acc2 falls back to one buffer; layer 2 of the next tile is issued between the two passes).
// Roles and the operand producers are those of sa_fused_tc.cu (tc_producer.cuh).
#include "tc_producer.cuh"

namespace {
using namespace tc;

constexpr int BM = kBM;
constexpr int BK = kBK;
constexpr int kEpiW = 8;
constexpr int kFirstEpiWarp = kProdWarps;                 // producers: warps 0 .. 15
constexpr int kMetaWarp = kProdWarps + kEpiW;             // 24
constexpr int kMmaWarp = kMetaWarp + 1;                   // 25 (highest id: first pick of its scheduler)
constexpr int kThreads = (kProdWarps + kEpiW + 2) * 32;   // 26 warps
constexpr int kMaxStages = 4;
constexpr int kABytes = kTileBytes;
constexpr int kC3 = 128;                                  // output channels per layer-3 pass = UMMA M

struct SatParams {
    const float *h; int ldh; int c1;
    const int32_t *idx; const float *xyz; const float *centres; const float *wxyz;
    int n, m, ns;
    long long rows, tiles;
    const uint8_t *w2blob; int n2, nkb1;          // layer 2: N = n2 = c2, K-blocks of c1 (fused.pack_tc image)
    const uint32_t *w3hi; const uint32_t *w3lo;   // layer 3: (128, c2 / 2) bf16 pairs, row = output channel
    const float *b2; const float *b3; int c2, nkb2;
    int c3, nm3;                                  // output channels, passes of 128 channels (1 or 2)
    int nb2;                                      // layer-2 accumulator buffers (2 when TMEM has room)
    float *y; int ldy;
    int stages;
    // duplicate-skipping mode (group_compact.cu): only the unique rows of every group are pushed through the MLP.
    // cmap[u] = centre of compact row u, jmap[u] = its neighbour index, *rows_dev = number of compact rows (device).
    const int32_t *cmap; const int32_t *jmap; const long long *rows_dev;
    int cmap_align;                               // 1, or 8: every aligned group of eight list rows has one centre
    unsigned long long *prof;                     // optional stopwatch buffer (32 u64 per CTA, tools/prof_sat.py) or nullptr
    int dbg;                                      // stopwatch build only: bit0 = the pooling epilogue only releases the accumulator (garbage results)
};

struct SmemLayout {
    uint32_t off_w2, off_a2, off_ring, off_meta, off_wx, off_bias, off_cid, off_bars, off_tmem, total;
};
__host__ __device__ inline SmemLayout make_layout(const SatParams &p, int stages) {
    SmemLayout L;
    uint32_t o = 0;
    L.off_w2 = o;   o += (uint32_t)p.nkb1 * 2u * p.n2 * 128u;
    L.off_a2 = o;   o += (uint32_t)p.nkb2 * 2u * kABytes;        // per K-block: hi tile | lo tile
    L.off_ring = o; o += (uint32_t)stages * 2u * kABytes;
    L.off_meta = o; o += kMetaDepth * BM * sizeof(RowMeta);
    L.off_wx = o;   o += 3u * p.nkb1 * BK * 4u;
    L.off_bias = o; o += (128 + 256) * 4;
    L.off_cid = o;  o += kEpiW * 64 * 4;                          // compact mode: centre ids of each epilogue warp's 64 columns
    L.off_bars = o; o += (2 * kMaxStages + 12 + 2 * kMetaDepth) * 8;
    L.off_tmem = o; o += 16;
    L.total = o;
    return L;
}

// Row metadata in compact mode: row u of the compact list is (centre cmap[u], neighbour jmap[u]); same software
// pipelining as tc::meta_run (the list entries of the NEXT tile are in flight while this tile's coordinates arrive).
__device__ __forceinline__ void meta_run_compact(const ProducerArgs &a, int lane, const int32_t *__restrict__ cmap,
                                                 const int32_t *__restrict__ jmap) {
    const long long first = blockIdx.x, stride = gridDim.x;
    const uint32_t m = (uint32_t)a.m;
    const long long last_row = a.rows - 1;
    constexpr int Q = kBM / 32;
    int cn[Q], jn[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { cn[q] = 0; jn[q] = 0; }
    if (first < a.items) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const long long u = min(first * kBM + q * 32 + lane, last_row);
            cn[q] = __ldg(cmap + u);
            jn[q] = __ldg(jmap + u);
        }
    }
    long long it = 0;
    for (long long item = first; item < a.items; item += stride, ++it) {
        const int slot = (int)(it % kMetaDepth);
        const long long row0 = item * kBM;
        int src[Q];
        float pj[Q][3], pc[Q][3];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const bool live = row0 + q * 32 + lane <= last_row;
            const uint32_t centre = (uint32_t)cn[q];
            const long long s = (long long)(centre / m) * a.n + jn[q];
            src[q] = live ? (int)s : 0;
            const float *gj = a.xyz + s * 3, *gc = a.centres + (long long)centre * 3;
            pj[q][0] = __ldg(gj); pj[q][1] = __ldg(gj + 1); pj[q][2] = __ldg(gj + 2);
            pc[q][0] = __ldg(gc); pc[q][1] = __ldg(gc + 1); pc[q][2] = __ldg(gc + 2);
            if (!live) { pj[q][0] = pc[q][0]; pj[q][1] = pc[q][1]; pj[q][2] = pc[q][2]; }
        }
        if (item + stride < a.items) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const long long u = min((item + stride) * kBM + q * 32 + lane, last_row);
                cn[q] = __ldg(cmap + u);
                jn[q] = __ldg(jmap + u);
            }
        }
        mbar_wait(&a.meta_empty[slot], (uint32_t)((it / kMetaDepth) & 1) ^ 1);
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
}

// maximum of sixteen accumulator values in eight 3-input FMNMX3 (sm_100) instead of fifteen FMNMX
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float max16(const uint32_t (&v)[16]) {
    const float a = max3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    const float b = max3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    const float c = max3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
    const float d = max3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
    const float e = max3(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
    return fmaxf(max3(a, b, c), max3(d, e, __uint_as_float(v[15])));
}

// one run of equal centre ids ends: combine its maximum into the pooled output (rare: out of line, one copy of the code)
__device__ __noinline__ void flush_run(float *dst, float v) {
    atomicMax(reinterpret_cast<unsigned int *>(dst), __float_as_uint(v));
}

// COMPACT (duplicate-skipping rows) is a kernel template parameter, so the dense instantiation carries none of its code:
// 0 = dense groups of nsample rows, 1 = compact row list with runs of any length, 8 = compact list whose groups are
// padded to multiples of eight rows (pn2_group_unique_count_i32 with align 8): pooled eight columns at a time
template <bool FAST, int COMPACT, bool PROF>
__global__ void __maxnreg__(72) sa_fused_t_tc_kernel(const SatParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (pn2_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = p.nkb1 * BK;
    const SmemLayout L = make_layout(p, p.stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint64_t *empty = full + kMaxStages;
    uint64_t *acc2_full = empty + kMaxStages;   // [2]
    uint64_t *acc2_empty = acc2_full + 2;       // [2]
    uint64_t *a2_full = acc2_empty + 2;         // [2]: per K-block of the layer-2 activation tile A2
    uint64_t *a2_empty = a2_full + 2;           // [2]
    uint64_t *acc3_full = a2_empty + 2;
    uint64_t *acc3_empty = acc3_full + 1;
    uint64_t *w_full = acc3_empty + 1;          // W2 in shared memory (bulk copy)
    uint64_t *w3_full = w_full + 1;             // W3 in tensor memory (epilogue warps)
    uint64_t *meta_full = w3_full + 1;
    uint64_t *meta_empty = meta_full + kMetaDepth;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_tmem);

    const uint32_t w2_bytes = (uint32_t)p.nkb1 * 2u * p.n2 * 128u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], kGroupWarps);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc2_full[a], 1);
            mbar_init(&acc2_empty[a], kEpiW);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a2_full[a], kEpiW);
            mbar_init(&a2_empty[a], 1);
        }
        mbar_init(acc3_full, 1);
        mbar_init(acc3_empty, kEpiW);
        mbar_init(w_full, 1);
        mbar_init(w3_full, kEpiW);
        for (int q = 0; q < kMetaDepth; ++q) {
            mbar_init(&meta_full[q], 1);
            mbar_init(&meta_empty[q], kProdWarps);
        }
        fence_barrier_init();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pn2_smem_u32(w_full)), "r"(w2_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         pn2_smem_u32(smem + L.off_w2)),
                     "l"(p.w2blob), "r"(w2_bytes), "r"(pn2_smem_u32(w_full))
                     : "memory");
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
    {
        float *wxs = reinterpret_cast<float *>(smem + L.off_wx);
        for (int i = threadIdx.x; i < 3 * kpad; i += kThreads) {
            const int c = i / kpad, k = i % kpad;
            wxs[i] = k < p.c1 ? __ldg(p.wxyz + c * p.c1 + k) : 0.f;
        }
        float *bias_s = reinterpret_cast<float *>(smem + L.off_bias);
        for (int i = threadIdx.x; i < 128 + p.nm3 * kC3; i += kThreads)
            bias_s[i] = i < 128 ? (i < p.c2 ? __ldg(p.b2 + i) : 0.f) : (i - 128 < p.c3 ? __ldg(p.b3 + (i - 128)) : 0.f);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t col_acc2 = 0;                                   // buffer b at b * n2
    const uint32_t col_w3 = (uint32_t)(p.nb2 * p.n2);              // pass j: hi at + j * c2, lo c2 / 2 columns after it
    const uint32_t col_acc3 = col_w3 + (uint32_t)(p.nm3 * p.c2);
    const int half_c2 = p.c2 >> 1;

    // compact mode: the row count is data dependent and lives on the device (no host synchronisation anywhere)
    constexpr bool compact = COMPACT != 0;
    const long long n_rows = compact ? *p.rows_dev : p.rows;
    const long long n_tiles = compact ? (n_rows + BM - 1) / BM : p.tiles;
    const long long first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < n_tiles ? (int)((n_tiles - first + stride - 1) / stride) : 0;

    ProducerArgs pa;
    pa.x = p.h; pa.ldx = p.ldh; pa.cin = p.c1; pa.rows = n_rows;
    pa.x2 = nullptr; pa.ldx2 = 0; pa.kb_split = 0x7fffffff;
    pa.vec_ok = ((p.ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.h) & 15) == 0);
    pa.idx = p.idx; pa.xyz = p.xyz; pa.centres = p.centres; pa.n = p.n; pa.m = p.m; pa.ns = p.ns;
    pa.nkb = p.nkb1; pa.stages = p.stages; pa.nchunks = 1; pa.items = n_tiles;
    pa.ring = smem + L.off_ring; pa.stage_bytes = 2 * kABytes; pa.full = full; pa.empty = empty;
    pa.meta = reinterpret_cast<RowMeta *>(smem + L.off_meta); pa.meta_full = meta_full; pa.meta_empty = meta_empty;
    pa.wxs = reinterpret_cast<const float *>(smem + L.off_wx); pa.kpad = kpad; pa.prof = p.prof;

    if (warp < kProdWarps) {
        // =============================== producers (tc_producer.cuh) ===============================
        producer_run<true, FAST, PROF>(pa, (int)threadIdx.x, [](long long, int, int) {});
    } else if (warp == kMetaWarp) {
        if (compact) meta_run_compact(pa, lane, p.cmap, p.jmap);
        else meta_run<PROF>(pa, lane);
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer ===============================
        // All 32 lanes run the loops (uniform operands, tc::elect_one()); one elected lane issues.
        // Tensor-pipe order  M2(0) M2(1) | M2(2).kb0 M3(0) M2(2).kb1 | M2(3).kb0 M3(1) M2(3).kb1 | ...   (M3(0) M2(1) M3(0)' | ... with one acc2 buffer)
        // (a CTA without a tile -- possible in compact mode, where the grid is sized for the upper bound -- still has
        // to see its weight copies land before it may exit)
        mbar_wait(w_full, 0);
        mbar_wait(w3_full, 0);
        tc_fence_after_sync();
        if (my_tiles > 0) {
            const uint32_t idesc2 = make_idesc_bf16(BM, p.n2);
            const uint32_t idesc3 = make_idesc_bf16(kC3, BM);       // M = channels, N = the tile's 128 rows
            const uint32_t w2a = pn2_smem_u32(smem + L.off_w2), a2a = pn2_smem_u32(smem + L.off_a2);
            int stage = 0;
            uint32_t phase = 0;
            unsigned long long w_ring = 0, w_acc2 = 0, w_a2 = 0, w_acc3 = 0;
            const long long t_begin = PROF ? clock64() : 0;
            // K-blocks [kb_lo, kb_hi) of layer 2 of tile `it` (the ring is consumed in tile / K-block order)
            auto issue_m2 = [&](int it, int kb_lo, int kb_hi) {
                const int buf = p.nb2 == 2 ? (it & 1) : 0;
                const int use = p.nb2 == 2 ? (it >> 1) : it;
                if (kb_lo == 0) {
                    mbar_wait_timed<PROF>(&acc2_empty[buf], (uint32_t)(use & 1) ^ 1, w_acc2);
                    tc_fence_after_sync();
                }
                const uint32_t d = tmem_base + col_acc2 + (uint32_t)(buf * p.n2);
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait_timed<PROF>(&full[stage], phase, w_ring);
                    tc_fence_after_sync();
                    const uint32_t sa = pn2_smem_u32(smem + L.off_ring + (size_t)stage * 2 * kABytes);
                    const uint32_t a_hi = desc_lo(sa), a_lo = desc_lo(sa + kABytes);
                    const uint32_t wb = w2a + (uint32_t)kb * 2u * p.n2 * 128u;
                    const uint32_t b_hi = desc_lo(wb), b_lo = desc_lo(wb + p.n2 * 128u);
                    const int krem = p.c1 - kb * BK;
                    const int ksteps = krem >= BK ? 4 : (krem + 15) >> 4;
                    if (elect_one()) {
                        if (ksteps == 4 && kb > 0) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                mma_ss_lo(d, a_hi + ks * 2, b_hi + ks * 2, idesc2, 1u);
                                mma_ss_lo(d, a_hi + ks * 2, b_lo + ks * 2, idesc2, 1u);
                                mma_ss_lo(d, a_lo + ks * 2, b_hi + ks * 2, idesc2, 1u);
                            }
                        } else {
                            for (int ks = 0; ks < ksteps; ++ks) {
                                mma_ss_lo(d, a_hi + ks * 2, b_hi + ks * 2, idesc2, (kb | ks) ? 1u : 0u);
                                mma_ss_lo(d, a_hi + ks * 2, b_lo + ks * 2, idesc2, 1u);
                                mma_ss_lo(d, a_lo + ks * 2, b_hi + ks * 2, idesc2, 1u);
                            }
                        }
                        mma_commit(&empty[stage]);
                        if (kb == p.nkb1 - 1) mma_commit(&acc2_full[buf]);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            };
            // one pass of layer 3 (transposed): acc3^T = W3[pass] (TMEM) x A2^T (shared memory, written by the epilogue)
            auto issue_m3 = [&](int it, int j) {
                // A2 is handed over K-block by K-block (64 channels of the layer-2 activation): layer 3 starts on the first
                // while the epilogue still converts the second, and -- what shortens the E2 -> M3 -> E2 loop of the
                // stopwatch -- the epilogue may overwrite a K-block as soon as the MMAs that read it have completed.
                const int cnt = it * p.nm3 + j;
                mbar_wait_timed<PROF>(acc3_empty, (uint32_t)(cnt & 1) ^ 1, w_acc3);
                const uint32_t d3 = tmem_base + col_acc3;
                const uint32_t w3h = tmem_base + col_w3 + (uint32_t)(j * p.c2), w3l = w3h + (uint32_t)half_c2;
                const int ksteps3 = p.c2 >> 4;
                for (int kb = 0; kb < p.nkb2; ++kb) {
                    if (j == 0) mbar_wait_timed<PROF>(&a2_full[kb], (uint32_t)(it & 1), w_a2);
                    tc_fence_after_sync();
                    const uint32_t tb = a2a + (uint32_t)kb * 2u * kABytes;
                    const int ks_end = min(ksteps3, kb * 4 + 4);
                    if (elect_one()) {
                        for (int ks = kb * 4; ks < ks_end; ++ks) {
                            const uint32_t b_hi = desc_lo(tb) + (uint32_t)((ks & 3) * 2);
                            const uint32_t b_lo = desc_lo(tb + kABytes) + (uint32_t)((ks & 3) * 2);
                            mma_ts_lo(d3, w3h + ks * 8, b_hi, idesc3, ks ? 1u : 0u);
                            mma_ts_lo(d3, w3h + ks * 8, b_lo, idesc3, 1u);
                            mma_ts_lo(d3, w3l + ks * 8, b_hi, idesc3, 1u);
                        }
                        if (kb == p.nkb2 - 1) mma_commit(acc3_full);
                        if (j == p.nm3 - 1) mma_commit(&a2_empty[kb]);
                    }
                    __syncwarp();
                }
            };
            issue_m2(0, 0, p.nkb1);
            if (p.nb2 == 2 && my_tiles > 1) issue_m2(1, 0, p.nkb1);
            for (int it = 0; it < my_tiles; ++it) {
                if (p.nb2 == 2) {
                    // double-buffered acc2: layer 2 runs TWO tiles ahead.  M3(it) and M2(it + 2) both become possible when the
                    // conversion epilogue E2(it) ends, but M3(it) also needs the layer-3 accumulator back from the pooling
                    // epilogue of tile it - 1, which the epilogue warps only start after E2(it): the tensor pipe idled
                    // through that wait (1.5 k of 7.1 k cycles per tile in the stopwatch).  The FIRST K-block of M2(it + 2)
                    // (about as long as the wait) goes in front of M3(it), the rest behind it; all of M2(it + 2) in front
                    // delayed M3(it) and with it the hand-back of A2 (measured: 7.7 k cycles per tile).
                    if (it + 2 < my_tiles) issue_m2(it + 2, 0, 1);
                    for (int j = 0; j < p.nm3; ++j) issue_m3(it, j);
                    if (it + 2 < my_tiles && p.nkb1 > 1) issue_m2(it + 2, 1, p.nkb1);
                } else {
                    // single acc2 (it is free once this tile's conversion is done, which pass 0 waits for anyway):
                    // layer 2 of the next tile sits between the passes and covers the pooling of pass 0
                    issue_m3(it, 0);
                    if (it + 1 < my_tiles) issue_m2(it + 1, 0, p.nkb1);
                    for (int j = 1; j < p.nm3; ++j) issue_m3(it, j);
                }
            }
            if (PROF && lane == 0) {
                unsigned long long *o = p.prof + (size_t)blockIdx.x * 32;
                o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_ring; o[2] = w_acc2; o[3] = w_a2; o[4] = w_acc3;
                o[5] = (unsigned long long)my_tiles;
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        // order  W3 -> TMEM | E2(0) | E2(1) E3(0) | E2(2) E3(1) | ... | E3(last)
        const float *bias2 = reinterpret_cast<const float *>(smem + L.off_bias);
        const float *bias3 = bias2 + 128;
        const int ew = warp - kFirstEpiWarp;
        const int q = ew & 3, half = ew >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int r = q * 32 + lane;                 // tile row (E2) / output channel (W3 load, E3) of this thread
        {
            // W3 -> tensor memory: warps with half 0 load the hi words of their 32 channels, half 1 the lo words
            for (int j = 0; j < p.nm3; ++j) {
                const uint32_t *src = (half ? p.w3lo : p.w3hi) + (size_t)(j * kC3 + r) * half_c2;
                const uint32_t dst = lane_addr + col_w3 + (uint32_t)(j * p.c2 + half * half_c2);
                for (int j0 = 0; j0 < half_c2; j0 += 8) {
                    const uint4 u0 = __ldg(reinterpret_cast<const uint4 *>(src + j0));
                    const uint4 u1 = __ldg(reinterpret_cast<const uint4 *>(src + j0 + 4));
                    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    tmem_st8(dst + j0, w);
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(w3_full);
        }
        uint8_t *a2s = smem + L.off_a2;
        // compact mode: run-start masks of this warp's 64 columns (bit l of run_e / run_o = column 2l / 2l+1 begins a
        // new centre), computed once per tile when the centre ids are staged
        uint32_t run_e = 0, run_o = 0;
        // compact mode: centre ids of this warp's 64 columns of tile `it` (-1 marks columns past the end of the row list).
        // The list lives in global memory (L2): the load is issued a whole tile before its values are needed, the first
        // version fetched it at the start of e3 and stalled the epilogue for the full latency every tile (1.9 k of 8.4 k
        // cycles per tile in the stopwatch, tools/prof_sat.py)
        auto cid_fetch = [&](int it) {
            const long long u0 = (first + (long long)it * stride) * BM + half * 64;
            const long long left = n_rows - u0;
            const int ncols = left >= 64 ? 64 : (left > 0 ? (int)left : 0);
            int2 t = make_int2(-1, -1);
            if (2 * lane < ncols) t = __ldg(reinterpret_cast<const int2 *>(p.cmap + u0) + lane);
            if (2 * lane + 1 >= ncols) t.y = -1;
            return t;
        };
        int2 cid_next = make_int2(-1, -1);
        if (compact && my_tiles > 0) cid_next = cid_fetch(0);
        unsigned long long w_acc2f = 0, w_a2e = 0, w_acc3f = 0, t_e2 = 0, t_e3 = 0;
        auto e3 = [&](int it, int j) {
            // acc3^T: lane = channel, columns = tile rows -> in-thread max over the nsample rows of each centre
            const long long tile = first + (long long)it * stride;
            int *cid_s = reinterpret_cast<int *>(smem + L.off_cid) + ew * 64;
            if (compact && j == 0) {
                // centre ids of this warp's 64 columns (fetched one tile ahead, see cid_fetch) -> shared memory and the
                // run-start masks, before the accumulator is awaited
                __syncwarp();
                const int2 t = cid_next;
                reinterpret_cast<int2 *>(cid_s)[lane] = t;
                int prev = __shfl_up_sync(0xffffffffu, t.y, 1);
                if (lane == 0) prev = -2;                       // column 0 always opens a run
                run_e = __ballot_sync(0xffffffffu, t.x != prev);
                run_o = __ballot_sync(0xffffffffu, t.y != t.x);
                __syncwarp();
                if (it + 1 < my_tiles) cid_next = cid_fetch(it + 1);   // in flight during this tile's pooling and the next E2
            }
            mbar_wait_timed<PROF>(acc3_full, (uint32_t)((it * p.nm3 + j) & 1), w_acc3f);
            const long long te0 = PROF ? clock64() : 0;
            tc_fence_after_sync();
            const uint32_t t3 = lane_addr + col_acc3 + (uint32_t)(half * 64);     // this warp: columns half*64 .. +63
            if (PROF && (p.dbg & 1)) {
                // what-if experiment (tools/prof_sat.py): a free pooling epilogue
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc3_empty);
            } else if constexpr (COMPACT == 8) {
                // compact rows in groups of eight (group_compact.cu, align 8): an aligned group of eight columns belongs to
                // one centre, so its maximum is four static 3-input maxima and the run logic (warp-uniform) runs once per
                // GROUP: bit 4g of run_e says whether column 8g opens a new centre.  Per warp and pass: 8 groups instead
                // of 64 per-column tests -- the per-column version spent 3.9 k cycles per tile here against 1.1 k for
                // the dense kernel (tools/prof_sat.py), for ~10 % more rows through the MLP.
                const int ch = j * kC3 + r;
                const float b = bias3[ch];
                float *ych = p.y + ch;
                int cur = -1;
                float run = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t va[16], vb[16];
                    tmem_ld16(t3 + c0, va);
                    tmem_ld16(t3 + c0 + 16, vb);
                    tmem_ld_wait();
                    if (c0 == 32) {        // the last loads are done: the accumulator may be overwritten
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc3_empty);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t *v = g < 2 ? va + 8 * g : vb + 8 * (g - 2);
                        const float m = max3(max3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2])),
                                             max3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5])),
                                             fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
                        if ((run_e >> ((c0 >> 1) + 4 * g)) & 1u) {
                            if (cur >= 0 && ch < p.c3) flush_run(ych + (long long)cur * p.ldy, fmaxf(run + b, 0.f));
                            cur = cid_s[c0 + 8 * g];
                            run = m;
                        } else {
                            run = fmaxf(run, m);
                        }
                    }
                }
                if (cur >= 0 && ch < p.c3) flush_run(ych + (long long)cur * p.ldy, fmaxf(run + b, 0.f));
            } else if constexpr (COMPACT == 1) {
                // compact rows: the columns of a centre are a run of equal ids (warp-uniform), of any length and possibly
                // continued in the next warp / tile -> running max, one atomicMax per run and channel (y is zeroed).
                // Runs are long (tens of columns), so the columns are taken eight at a time: a group without a run start
                // is four 3-input maxima; only groups that contain a start walk their columns one by one.
                const int ch = j * kC3 + r;
                const float b = bias3[ch];
                float *ych = p.y + ch;
                const uint32_t run_any = run_e | run_o;
                int cur = -1;
                float run = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    {
                        uint32_t va[16], vb[16];
                        tmem_ld16(t3 + c0, va);
                        tmem_ld16(t3 + c0 + 16, vb);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; ++k) { v[k] = va[k]; v[16 + k] = vb[k]; }
                    }
                    if (c0 == 32) {        // the last loads are done: the accumulator may be overwritten
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc3_empty);
                    }
                    const uint32_t any16 = (run_any >> (c0 >> 1)) & 0xffffu;     // 16 bit pairs = these 32 columns
                    const uint32_t e16 = (run_e >> (c0 >> 1)) & 0xffffu, o16 = (run_o >> (c0 >> 1)) & 0xffffu;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (((any16 >> (4 * g)) & 0xfu) == 0u) {
                            float m0, m1, m2;
                            asm("max.f32 %0, %1, %2, %3;" : "=f"(m0) : "f"(__uint_as_float(v[8 * g])), "f"(__uint_as_float(v[8 * g + 1])), "f"(__uint_as_float(v[8 * g + 2])));
                            asm("max.f32 %0, %1, %2, %3;" : "=f"(m1) : "f"(__uint_as_float(v[8 * g + 3])), "f"(__uint_as_float(v[8 * g + 4])), "f"(__uint_as_float(v[8 * g + 5])));
                            asm("max.f32 %0, %1, %2, %3;" : "=f"(m2) : "f"(__uint_as_float(v[8 * g + 6])), "f"(__uint_as_float(v[8 * g + 7])), "f"(run));
                            asm("max.f32 %0, %1, %2, %3;" : "=f"(run) : "f"(m0), "f"(m1), "f"(m2));
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int k = 8 * g + e;
                                const float x = __uint_as_float(v[k]);
                                const bool start = (((e & 1) ? o16 : e16) >> (k >> 1)) & 1u;
                                if (start) {
                                    if (cur >= 0 && ch < p.c3) flush_run(ych + (long long)cur * p.ldy, fmaxf(run + b, 0.f));
                                    cur = cid_s[c0 + k];
                                    run = x;
                                } else {
                                    run = fmaxf(run, x);
                                }
                            }
                        }
                    }
                }
                if (cur >= 0 && ch < p.c3) flush_run(ych + (long long)cur * p.ldy, fmaxf(run + b, 0.f));
            } else {
            float cm[4];
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t va[16], vb[16];
                tmem_ld16(t3 + jp * 32, va);
                tmem_ld16(t3 + jp * 32 + 16, vb);
                tmem_ld_wait();
                cm[2 * jp] = max16(va);
                cm[2 * jp + 1] = max16(vb);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc3_empty);      // the accumulator is in registers: layer 3 of the next tile may start
            const int ch = j * kC3 + r;
            const float b = bias3[ch];
            const long long row_base = tile * BM + half * 64;       // first tile row covered by this warp
            auto emit = [&](float mval, long long first_row, bool atomic) {
                if (first_row < p.rows && ch < p.c3) {
                    const float o = fmaxf(mval + b, 0.f);
                    float *dst = p.y + (first_row / p.ns) * p.ldy + ch;
                    if (atomic) atomicMax(reinterpret_cast<unsigned int *>(dst), __float_as_uint(o));
                    else *dst = o;
                }
            };
            if (p.ns == 16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) emit(cm[j], row_base + j * 16, false);
            } else if (p.ns == 32) {
                emit(fmaxf(cm[0], cm[1]), row_base, false);
                emit(fmaxf(cm[2], cm[3]), row_base + 32, false);
            } else {
                const float m4 = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
                emit(m4, p.ns == 64 ? row_base : tile * BM, p.ns != 64);   // nsample 128: two warps share a centre
            }
            }   // dense rows
            if (PROF) t_e3 += (unsigned long long)(clock64() - te0);
        };
        const bool pipelined = p.nb2 == 2;     // pooling of tile i-1 after the conversion of tile i (see the MMA order)
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = pipelined ? (it & 1) : 0;
            const int use = pipelined ? (it >> 1) : it;
            // ---- E2: acc2[buf] -> bias, ReLU, bf16 hi/lo -> shared-memory operand A2 of layer 3 ----
            mbar_wait_timed<PROF>(&acc2_full[buf], (uint32_t)(use & 1), w_acc2f);
            mbar_wait_timed<PROF>(&a2_empty[0], (uint32_t)(it & 1) ^ 1, w_a2e);
            const long long t20 = PROF ? clock64() : 0;
            tc_fence_after_sync();
            // one K-block of A2 (64 channels: one pass of the loop below over both warp halves) is complete
            auto kblock_done = [&](int kb) {
                fence_proxy_async_smem();      // generic-proxy stores -> visible to the MMA's async-proxy operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2_full[kb]);
                if (kb + 1 < p.nkb2) {
                    mbar_wait_timed<PROF>(&a2_empty[kb + 1], (uint32_t)(it & 1) ^ 1, w_a2e);
                    tc_fence_after_sync();
                }
            };
            const uint32_t t_acc2 = lane_addr + col_acc2 + (uint32_t)(buf * p.n2);
            auto convert = [&](const uint32_t (&v)[16], int c0) {
                // 16 consecutive k of row r: two 16-byte chunks of the row's 128-byte line in K-block c0 / 64
                uint4 hi[2], lo[2];
                const float4 *b4 = reinterpret_cast<const float4 *>(bias2 + c0);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 bb = b4[j4];
                    const float x0 = fmaxf(__uint_as_float(v[4 * j4 + 0]) + bb.x, 0.f);
                    const float x1 = fmaxf(__uint_as_float(v[4 * j4 + 1]) + bb.y, 0.f);
                    const float x2 = fmaxf(__uint_as_float(v[4 * j4 + 2]) + bb.z, 0.f);
                    const float x3 = fmaxf(__uint_as_float(v[4 * j4 + 3]) + bb.w, 0.f);
                    uint2 h2, l2;
                    split4(make_float4(x0, x1, x2, x3), h2, l2);
                    if (j4 & 1) { hi[j4 >> 1].z = h2.x; hi[j4 >> 1].w = h2.y; lo[j4 >> 1].z = l2.x; lo[j4 >> 1].w = l2.y; }
                    else        { hi[j4 >> 1].x = h2.x; hi[j4 >> 1].y = h2.y; lo[j4 >> 1].x = l2.x; lo[j4 >> 1].y = l2.y; }
                }
                uint8_t *tile_hi = a2s + (size_t)(c0 >> 6) * 2 * kABytes;
                const int ci = (c0 & 63) >> 3;
                *reinterpret_cast<uint4 *>(tile_hi + sw128_offset(r, ci)) = hi[0];
                *reinterpret_cast<uint4 *>(tile_hi + sw128_offset(r, ci + 1)) = hi[1];
                *reinterpret_cast<uint4 *>(tile_hi + kABytes + sw128_offset(r, ci)) = lo[0];
                *reinterpret_cast<uint4 *>(tile_hi + kABytes + sw128_offset(r, ci + 1)) = lo[1];
            };
            // K-block kb of A2 = channels 64 kb .. 64 kb + 63: this warp converts the 16-column chunks at half * 16 and
            // half * 16 + 32 of it (those below n2); EVERY warp reports every K-block, with or without columns of its own
            for (int kb = 0; kb < p.nkb2; ++kb) {
                const int ca = kb * 64 + half * 16;
                if (ca + 32 < p.n2) {
                    uint32_t va[16], vb[16];
                    tmem_ld16(t_acc2 + ca, va);
                    tmem_ld16(t_acc2 + ca + 32, vb);
                    tmem_ld_wait();
                    convert(va, ca);
                    convert(vb, ca + 32);
                } else if (ca < p.n2) {
                    uint32_t va[16];
                    tmem_ld16(t_acc2 + ca, va);
                    tmem_ld_wait();
                    convert(va, ca);
                }
                kblock_done(kb);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc2_empty[buf]);
            if (PROF) t_e2 += (unsigned long long)(clock64() - t20);
            if (!pipelined) {
                for (int j = 0; j < p.nm3; ++j) e3(it, j);
            } else if (it > 0) {
                for (int j = 0; j < p.nm3; ++j) e3(it - 1, j);
            }
        }
        if (pipelined && my_tiles > 0)
            for (int j = 0; j < p.nm3; ++j) e3(my_tiles - 1, j);
        if (PROF && ew == 0 && lane == 0) {
            unsigned long long *o = p.prof + (size_t)blockIdx.x * 32;
            o[8] = w_acc2f; o[9] = w_a2e; o[10] = w_acc3f; o[12] = t_e2; o[13] = t_e3;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

unsigned long long *g_prof = nullptr;
int g_dbg = 0;

}  // namespace

// what-if switches of the stopwatch build (tools/prof_sat.py; results are garbage): bit0 = free pooling epilogue
PN2_API void pn2_sa_fused_t_set_debug(int bits) { g_dbg = bits; }

// stopwatch buffer for tools/prof_sat.py: 32 u64 per CTA (device memory) or NULL to disable (never used by the product)
PN2_API void pn2_sa_fused_t_set_profile(void *buf) { g_prof = static_cast<unsigned long long *>(buf); }

// pn2_sa_fused_tc_f32 with the last layer transposed (see the header of this file).
//   w3hi / w3lo: (ceil(c3 / 128) * 128, c2 / 2) uint32 each (rows past c3 zero), W3 split into bf16 hi / lo, word j of row o = (W3[o][2j], W3[o][2j+1])
//   with the even k in the low half (fused.pack_w3t).
// Supported: c3 <= 128 (the accumulator is padded to 128 channel lanes, only c3 are stored) or 256 (two passes of layer 3 over the same activation tile; the layer-2 accumulator
// is then single-buffered), c2 a multiple of 16 and <= 128 (n2 == c2), ns in {16, 32, 64, 128}; nsample 128 combines
// the two halves of a centre with atomicMax, so y must be zero-filled for it (not for 16 / 32 / 64).
// Duplicate-skipping mode: cmap / jmap / rows_dev from pn2_group_compact_i32 (all three or none) and cmap_align = the align
// that list was built with (1 or 8; 8 selects the group-wise pooling epilogue).  y must then be zero-filled (runs of a
// centre are combined with atomicMax); without them every nsample-row group is processed.
PN2_API int pn2_sa_fused_t_tc_f32(const float *h, int ldh, const int32_t *idx, const float *xyz, const float *centres,
                                  const float *wxyz, const void *w2blob, int n2, int nkb1, const float *b2,
                                  const void *w3hi, const void *w3lo, const float *b3, float *y, int ldy, int clouds,
                                  int n, int m, int ns, int c1, int c2, int c3, const int32_t *cmap, const int32_t *jmap,
                                  const long long *rows_dev, int cmap_align, cudaStream_t stream) {
    if ((cmap != nullptr) != (jmap != nullptr) || (cmap != nullptr) != (rows_dev != nullptr) ||
        (cmap && (reinterpret_cast<uintptr_t>(cmap) & 15))) {
        pn2_set_last_error("pn2_sa_fused_t_tc_f32: cmap / jmap / rows_dev go together (cmap 16-byte aligned)");
        return PN2_ERR_INVALID;
    }
    if (cmap && cmap_align != 1 && cmap_align != 8) {
        pn2_set_last_error("pn2_sa_fused_t_tc_f32: cmap_align must be 1 or 8 (the align given to pn2_group_unique_count_i32)");
        return PN2_ERR_INVALID;
    }
    if (!h || !idx || !xyz || !centres || !wxyz || !w2blob || !w3hi || !w3lo || !b2 || !b3 || !y || clouds < 0 ||
        n <= 0 || m < 0 || c1 <= 0 || c2 <= 0 || ldh < c1 || ldy < c3 || (long long)clouds * n > 2147483647LL) {
        pn2_set_last_error("pn2_sa_fused_t_tc_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    const int nm3 = (c3 + kC3 - 1) / kC3;
    const int nb2 = (2 * n2 + nm3 * c2 + BM <= 512) ? 2 : 1;
    if (c3 <= 0 || (c3 > kC3 && c3 != 2 * kC3) || !(ns == 16 || ns == 32 || ns == 64 || ns == 128) || (c2 & 15) || c2 > 128 || n2 != c2 ||
        nkb1 * BK < c1 || nb2 * n2 + nm3 * c2 + BM > 512 || ((reinterpret_cast<uintptr_t>(w3hi) | reinterpret_cast<uintptr_t>(w3lo)) & 15)) {
        pn2_set_last_error("pn2_sa_fused_t_tc_f32: unsupported shape");
        return PN2_ERR_UNSUPPORTED;
    }
    SatParams p = {};
    p.h = h; p.ldh = ldh; p.c1 = c1; p.idx = idx; p.xyz = xyz; p.centres = centres; p.wxyz = wxyz;
    p.n = n; p.m = m; p.ns = ns;
    p.rows = (long long)clouds * m * ns;
    p.tiles = (p.rows + BM - 1) / BM;
    p.w2blob = static_cast<const uint8_t *>(w2blob); p.n2 = n2; p.nkb1 = nkb1;
    p.w3hi = static_cast<const uint32_t *>(w3hi); p.w3lo = static_cast<const uint32_t *>(w3lo);
    p.b2 = b2; p.b3 = b3; p.c2 = c2; p.nkb2 = (c2 + BK - 1) / BK; p.y = y; p.ldy = ldy;
    p.c3 = c3; p.nm3 = nm3; p.nb2 = nb2;
    p.cmap = cmap; p.jmap = jmap; p.rows_dev = rows_dev; p.cmap_align = cmap_align;
    if (p.rows == 0) return PN2_OK;
    int stages = kMaxStages;
    SmemLayout L = make_layout(p, stages);
    while (stages > 2 && L.total + 1024 > 227 * 1024) L = make_layout(p, --stages);
    if (L.total + 1024 > 227 * 1024) {
        pn2_set_last_error("pn2_sa_fused_t_tc_f32: does not fit in shared memory");
        return PN2_ERR_UNSUPPORTED;
    }
    p.stages = stages;
    p.prof = g_prof;
    p.dbg = g_dbg;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)tc::persistent_grid(p.tiles, sms);
    const int vec_ok = ((ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(h) & 15) == 0);
    const size_t smem_bytes = L.total + 1024;
    const bool fast = tc::producer_fast(vec_ok, c1);
    const int mode = cmap ? cmap_align : 0;
    auto go = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
    };
    if (p.prof && fast) {        // stopwatch build (tools/prof_sat.py only)
        if (mode == 8) go(sa_fused_t_tc_kernel<true, 8, true>);
        else if (mode == 1) go(sa_fused_t_tc_kernel<true, 1, true>);
        else go(sa_fused_t_tc_kernel<true, 0, true>);
    } else if (mode == 8) {
        if (fast) go(sa_fused_t_tc_kernel<true, 8, false>);
        else go(sa_fused_t_tc_kernel<false, 8, false>);
    } else if (mode == 1) {
        if (fast) go(sa_fused_t_tc_kernel<true, 1, false>);
        else go(sa_fused_t_tc_kernel<false, 1, false>);
    } else {
        if (fast) go(sa_fused_t_tc_kernel<true, 0, false>);
        else go(sa_fused_t_tc_kernel<false, 0, false>);
    }
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
