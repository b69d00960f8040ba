// rcnn_front_tc.cu -- the RCNN input chain as ONE persistent tcgen05 kernel, sm_100a.
//
// Replaces, for the pooled ROI points (rows = B * 100 * 512 = 819 200 at batch 16), the reference chain
//   xyz_up_layer      SharedMLP [5 -> 128 -> 128]              (lib/net/rcnn_net.py:41-47, :168-171)
//   merge_down_layer  SharedMLP [256 -> 128] on cat[xyz_feature, rpn_feature]        (rcnn_net.py:174-176)
//   SA1 layer 1, per-point half   H = W1f . merged + b1        (pointnet2_modules.py:38-44 through QueryAndGroup; the
//                                 pair-wise half W1x . (x_j - centre) and the ReLU are applied by the SA kernel)
// Layer by layer each stage wrote a (rows x 128) fp32 activation (0.42 GB) that the next one read straight back: three
// HBM-bound launches moving ~3.5 GB.  Here a tile of 128 pooled rows is read once (5 + 128 floats per row) and only H
// is written: 0.44 GB in, 0.42 GB out.
//
// Per tile, all contractions BF16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulate in tensor memory):
//   producers (16 warps)  A1 = relu(Wpre . x[0:5] + bpre) in packed fp32 -> bf16 hi/lo -> swizzled smem ring (2 K-blocks)
//                         F  = the 128 rpn features of the row           -> bf16 hi/lo -> the same ring      (2 K-blocks)
//   weight warp           streams the eight 32 KB weight K-blocks a tile needs (pre-split, pre-swizzled by fused.pack_tc,
//                         L2-resident) through a second ring with cp.async.bulk, in the MMA warp's consumption order
//   MMA warp              M1  acc1  = A1 x Wup2^T                     (SS)
//                         M2F acc2  = F  x Wmerge[:, 128:256]^T       (SS)   -- does not wait for the epilogue
//                         M2X acc2 += EA x Wmerge[:, 0:128]^T         (TS: A operand in tensor memory)
//                         M3  acc3  = EA x W1f^T                      (TS)
//   epilogue (8 warps)    E1  acc1 -> +b, ReLU -> bf16 hi/lo -> EA (tensor memory, tcgen05.st)       [xyz_feature]
//                         E2  acc2 -> +b, ReLU -> bf16 hi/lo -> EA                                   [merged]
//                         E3  acc3 -> +b1 -> H rows (coalesced through a shared-memory transpose)
// TMEM: acc1 | acc2 | acc3 | EA = 4 x 128 columns.  EA is single: the chain E1 -> M2X -> E2 -> M3 of a tile is serial, so
// the tensor pipe is kept busy across tiles instead -- issue order  M2X(i) M1(i+1) M2F(i+1) M3(i), epilogue order
// E2(i) E1(i+1) E3(i).  Shared memory: 3 x 32 KB operand ring + 3 x 32 KB weight ring.
#include "tc_producer.cuh"
#include "tc_epilogue.cuh"

namespace {
using namespace tc;

constexpr int BM = kBM;
constexpr int BK = kBK;
constexpr int kC = 128;                                   // width of every layer of the chain
constexpr int kPre = 5;                                   // x, y, z, mask, depth
constexpr int kFirstEpiWarp = kProdWarps;                 // producers: warps 0 .. 15
constexpr int kWeightWarp = kProdWarps + kEpiWarps;       // 24
constexpr int kMmaWarp = kWeightWarp + 1;                 // 25 (highest id: first pick of its scheduler)
constexpr int kThreads = (kProdWarps + kEpiWarps + 2) * 32;
constexpr int kStageBytes = 2 * kTileBytes;               // hi tile | lo tile = 32 KB (operands and weights alike)
constexpr int kSA = 3, kSW = 3;
constexpr uint32_t kColAcc1 = 0, kColAcc2 = 128, kColAcc3 = 256, kColEA = 384;

struct FrontParams {
    const float *x; int ldx; int off_f; long long rows, tiles;
    const float *wpre;                     // (kPre + 1, 128): the weight rows of xyz_up layer 1 (input-major), then its bias
    const uint8_t *w_up2, *w_merge, *w_sa; // fused.pack_tc images: 2, 4, 2 K-blocks of 32 KB
    const float *b_up2, *b_merge, *b_sa;
    float *h; int ldh;
    unsigned long long *prof;              // optional stopwatch buffer (32 u64 per CTA, tools/prof_front.py) or nullptr
    int mode;                              // tuning experiments (tools/prof_front.py): bit1 = M2F issued before M3
};

struct Smem {
    uint32_t off_a, off_w, off_wpre, off_xs, off_bias, off_bars, off_tmem, off_clk, total;
};
__host__ __device__ inline Smem make_layout() {
    Smem L;
    uint32_t o = 0;
    L.off_a = o;    o += kSA * kStageBytes;
    L.off_w = o;    o += kSW * kStageBytes;
    L.off_wpre = o; o += (kPre + 1) * kC * 4;
    L.off_xs = o;   o += kGroups * 2 * BM * 8 * 4;     // row heads (x y z mask depth + pad) of the current and next tile, per group
    L.off_bias = o; o += 3 * kC * 4;
    L.off_bars = o; o += (2 * kSA + 2 * kSW + 6 + 2) * 8;
    L.off_tmem = o; o += 16;
    L.off_clk = o;  o += 8 * 8;        // stopwatch: clock of each epilogue warp's last EA hand-off (PROF only)
    L.total = o;
    return L;
}

template <bool PROF>
__global__ void __maxnreg__(72) rcnn_front_tc_kernel(const FrontParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (pn2_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Smem L = make_layout();
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint64_t *a_empty = a_full + kSA;
    uint64_t *w_full = a_empty + kSA;
    uint64_t *w_empty = w_full + kSW;
    uint64_t *acc_full = w_empty + kSW;      // [3]
    uint64_t *acc_empty = acc_full + 3;      // [3]
    uint64_t *ea_full = acc_empty + 3;
    uint64_t *ea_empty = ea_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_tmem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSA; ++s) { mbar_init(&a_full[s], kGroupWarps); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kSW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 3; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], kEpiWarps); }
        mbar_init(ea_full, kEpiWarps);
        mbar_init(ea_empty, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
    {
        float *wp = reinterpret_cast<float *>(smem + L.off_wpre);
        for (int i = threadIdx.x; i < (kPre + 1) * kC; i += kThreads) wp[i] = __ldg(p.wpre + i);
        float *bs = reinterpret_cast<float *>(smem + L.off_bias);
        for (int i = threadIdx.x; i < 3 * kC; i += kThreads)
            bs[i] = __ldg((i < kC ? p.b_up2 : i < 2 * kC ? p.b_merge : p.b_sa) + (i & (kC - 1)));
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const long long first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < p.tiles ? (int)((p.tiles - first + stride - 1) / stride) : 0;

    if (warp < kProdWarps) {
        // =============================== producers ===============================
        // four operand steps per tile: [A1 k 0..63, A1 k 64..127, F k 0..63, F k 64..127]; group g of eight warps takes
        // steps g and g + 2, i.e. one computed K-block and one loaded K-block of the same tile.
        // Latency plan (the first version re-read the five inputs of a row from global memory inside the FMA loop and
        // spilled its weights: 13 k of its 17 k cycles per tile were exposed load latency): a group stages the 32-byte
        // head of the tile's 128 rows in shared memory, ONE tile ahead (one 16-byte load per thread, in flight during the
        // whole previous tile); the eight feature loads of a thread are issued before the computed step and consumed
        // after it; the first-layer weights are re-read from shared memory per pass instead of pinning 24 registers.
        const int ptid = threadIdx.x, pw = ptid >> 5;
        const int group = pw % kGroups, wg = pw / kGroups;
        const int gt = wg * 32 + lane;                     // thread index inside the group, 0 .. 255
        const int rsub = wg * 2 + (lane >> 4);
        const int kq = (lane & 15) * 4;
        const uint32_t toff = (uint32_t)(((rsub >> 3) << 10) + ((rsub & 7) << 7)) +
                              ((((uint32_t)((lane & 15) >> 1) ^ (uint32_t)(rsub & 7)) << 4) | ((uint32_t)(lane & 1) << 3));
        const int k = group * BK + kq;                     // this thread's four channels, in both of its steps
        // the thread's first-layer weights (5 rows + bias, 4 channels each) stay in registers: re-reading them from
        // shared memory per pass cost 3.8 k shared-memory wavefronts per tile on a port the MMA operands already fill
        float2 w01[kPre + 1], w23[kPre + 1];
        {
            const float *wk = reinterpret_cast<const float *>(smem + L.off_wpre) + k;
#pragma unroll
            for (int c = 0; c <= kPre; ++c) {
                const float4 w = *reinterpret_cast<const float4 *>(wk + c * kC);
                w01[c] = make_float2(w.x, w.y);
                w23[c] = make_float2(w.z, w.w);
            }
        }
        float *xs = reinterpret_cast<float *>(smem + L.off_xs) + group * (2 * BM * 8);     // [2][128 rows][8 floats]
        int stage = group;                                 // step 4 * it + group  -> stage (4 it + group) % 3
        uint32_t phase = 0;
        auto advance2 = [&]() {                            // two steps further in the ring
            stage += 2;
            if (stage >= kSA) { stage -= kSA; phase ^= 1; }
        };
        auto head_load = [&](int it) {                     // this thread's 16 bytes of the row heads of tile `it`
            const long long row = min((first + (long long)it * stride) * BM + (gt >> 1), p.rows - 1);
            return __ldg(reinterpret_cast<const float4 *>(p.x + row * p.ldx + (gt & 1) * 4));
        };
        unsigned long long w_stage = 0, t_bar = 0, t_comp = 0, t_fst = 0;
        const long long t_begin = PROF ? clock64() : 0;
        float4 xhead = make_float4(0.f, 0.f, 0.f, 0.f);
        if (my_tiles > 0) xhead = head_load(0);
        for (int it = 0; it < my_tiles; ++it) {
            const long long row0 = (first + (long long)it * stride) * BM + rsub;
            float *xb = xs + (it & 1) * (BM * 8);
            *reinterpret_cast<float4 *>(xb + gt * 4) = xhead;          // row gt / 2, floats (gt & 1) * 4 ..
            if (it + 1 < my_tiles) xhead = head_load(it + 1);
            float4 freg[kPasses];
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
                const long long row = min(row0 + ps * kRowsPerPass, p.rows - 1);
                freg[ps] = __ldg(reinterpret_cast<const float4 *>(p.x + row * p.ldx + p.off_f + k));
            }
            const long long tb0 = PROF ? clock64() : 0;
            named_bar_sync(1 + group, kGroupWarps * 32);                // the group's row heads are in shared memory
            if (PROF) t_bar += (unsigned long long)(clock64() - tb0);
            // ---- computed step: relu(bpre + Wpre . x[0:5]) for channels k .. k + 3 ----
            {
                uint8_t *sbase = smem + L.off_a + (size_t)stage * kStageBytes + toff;
                mbar_wait_lazy_timed<PROF>(&a_empty[stage], phase ^ 1, w_stage);
                const long long tc0 = PROF ? clock64() : 0;
#pragma unroll
                for (int ps = 0; ps < kPasses; ++ps) {
                    const float *xr = xb + (ps * kRowsPerPass + rsub) * 8;
                    float4 q0;
                    lds128(xr, q0);
                    const float q4 = xr[4];
                    const float in[kPre] = {q0.x, q0.y, q0.z, q0.w, q4};
                    float2 t01 = w01[kPre], t23 = w23[kPre];
#pragma unroll
                    for (int c = 0; c < kPre; ++c) {
                        const float2 xc = make_float2(in[c], in[c]);
                        t01 = __ffma2_rn(w01[c], xc, t01);
                        t23 = __ffma2_rn(w23[c], xc, t23);
                    }
                    const float4 v = make_float4(fmaxf(t01.x, 0.f), fmaxf(t01.y, 0.f), fmaxf(t23.x, 0.f), fmaxf(t23.y, 0.f));
                    uint2 hi, lo;
                    split4(v, hi, lo);
                    *reinterpret_cast<uint2 *>(sbase + ps * 2048) = hi;
                    *reinterpret_cast<uint2 *>(sbase + kTileBytes + ps * 2048) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
                advance2();
                if (PROF) t_comp += (unsigned long long)(clock64() - tc0);
            }
            // ---- loaded step: the row's rpn features ----
            {
                uint8_t *sbase = smem + L.off_a + (size_t)stage * kStageBytes + toff;
                mbar_wait_lazy_timed<PROF>(&a_empty[stage], phase ^ 1, w_stage);
                const long long tf0 = PROF ? clock64() : 0;
#pragma unroll
                for (int ps = 0; ps < kPasses; ++ps) {
                    uint2 hi, lo;
                    split4(freg[ps], hi, lo);
                    *reinterpret_cast<uint2 *>(sbase + ps * 2048) = hi;
                    *reinterpret_cast<uint2 *>(sbase + kTileBytes + ps * 2048) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
                advance2();
                if (PROF) t_fst += (unsigned long long)(clock64() - tf0);
            }
        }
        if (PROF && ptid == 0) {
            unsigned long long *o = p.prof + (size_t)blockIdx.x * 32;
            o[16] = (unsigned long long)(clock64() - t_begin); o[17] = w_stage; o[18] = t_bar; o[19] = t_comp; o[20] = t_fst;
        }
    } else if (warp == kWeightWarp) {
        // =============================== weight stream ===============================
        // consumption order of the MMA warp:  U0 U1 F0 F1 | per tile: X0 X1, (next tile's U0 U1 F0 F1), S0 S1
        if (lane == 0 && my_tiles > 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto push = [&](const uint8_t *src) {
                mbar_wait_lazy(&w_empty[stage], phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pn2_smem_u32(&w_full[stage])),
                             "r"((uint32_t)kStageBytes)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 pn2_smem_u32(smem + L.off_w + (size_t)stage * kStageBytes)),
                             "l"(src), "r"((uint32_t)kStageBytes), "r"(pn2_smem_u32(&w_full[stage]))
                             : "memory");
                if (++stage == kSW) { stage = 0; phase ^= 1; }
            };
            auto push_head = [&]() {
                push(p.w_up2);
                push(p.w_up2 + kStageBytes);
                push(p.w_merge + 2 * kStageBytes);
                push(p.w_merge + 3 * kStageBytes);
            };
            push_head();
            for (int it = 0; it < my_tiles; ++it) {
                push(p.w_merge);
                push(p.w_merge + kStageBytes);
                const bool more = it + 1 < my_tiles;
                if (more) { push(p.w_up2); push(p.w_up2 + kStageBytes); }
                if (more && (p.mode & 2)) { push(p.w_merge + 2 * kStageBytes); push(p.w_merge + 3 * kStageBytes); }
                push(p.w_sa);
                push(p.w_sa + kStageBytes);
                if (more && !(p.mode & 2)) { push(p.w_merge + 2 * kStageBytes); push(p.w_merge + 3 * kStageBytes); }
            }
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer ===============================
        if (my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(BM, kC);
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            uint32_t ea_use = 0;
            unsigned long long w_a = 0, w_w = 0, w_ea2 = 0, w_ea3 = 0, w_acc = 0, wake = 0;
            const long long t_begin = PROF ? clock64() : 0;
            // one K-block with both operands in shared memory
            auto step_ss = [&](uint32_t d, bool fresh) {
                mbar_wait_timed<PROF>(&a_full[sa], pa, w_a);
                mbar_wait_timed<PROF>(&w_full[sw], pw, w_w);
                tc_fence_after_sync();
                const uint32_t aa = pn2_smem_u32(smem + L.off_a + (size_t)sa * kStageBytes);
                const uint32_t wa = pn2_smem_u32(smem + L.off_w + (size_t)sw * kStageBytes);
                const uint32_t a_hi = desc_lo(aa), a_lo = desc_lo(aa + kTileBytes);
                const uint32_t b_hi = desc_lo(wa), b_lo = desc_lo(wa + kTileBytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        mma_ss_lo(d, a_hi + ks * 2, b_hi + ks * 2, idesc, (fresh && ks == 0) ? 0u : 1u);
                        mma_ss_lo(d, a_hi + ks * 2, b_lo + ks * 2, idesc, 1u);
                        mma_ss_lo(d, a_lo + ks * 2, b_hi + ks * 2, idesc, 1u);
                    }
                    mma_commit(&a_empty[sa]);
                    mma_commit(&w_empty[sw]);
                }
                __syncwarp();
                if (++sa == kSA) { sa = 0; pa ^= 1; }
                if (++sw == kSW) { sw = 0; pw ^= 1; }
            };
            // one K-block with the A operand (EA, written by the epilogue) in tensor memory: kb = 0 / 1 -> k 0..63 / 64..127
            auto step_ts = [&](uint32_t d, int kb, bool fresh) {
                mbar_wait_timed<PROF>(&w_full[sw], pw, w_w);
                tc_fence_after_sync();
                const uint32_t wa = pn2_smem_u32(smem + L.off_w + (size_t)sw * kStageBytes);
                const uint32_t b_hi = desc_lo(wa), b_lo = desc_lo(wa + kTileBytes);
                const uint32_t e_hi = tmem_base + kColEA + (uint32_t)(kb * 32), e_lo = e_hi + 64;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        mma_ts_lo(d, e_hi + ks * 8, b_hi + ks * 2, idesc, (fresh && ks == 0) ? 0u : 1u);
                        mma_ts_lo(d, e_hi + ks * 8, b_lo + ks * 2, idesc, 1u);
                        mma_ts_lo(d, e_lo + ks * 8, b_hi + ks * 2, idesc, 1u);
                    }
                    mma_commit(&w_empty[sw]);
                }
                __syncwarp();
                if (++sw == kSW) { sw = 0; pw ^= 1; }
            };
            auto head_m1 = [&](int it) {       // M1(it)
                mbar_wait_timed<PROF>(&acc_empty[0], (uint32_t)(it & 1) ^ 1, w_acc);
                tc_fence_after_sync();
                step_ss(tmem_base + kColAcc1, true);
                step_ss(tmem_base + kColAcc1, false);
                if (elect_one()) mma_commit(&acc_full[0]);
                __syncwarp();
            };
            auto head_m2f = [&](int it) {      // M2F(it): the rpn-feature half of merge_down, no dependence on the epilogue
                mbar_wait_timed<PROF>(&acc_empty[1], (uint32_t)(it & 1) ^ 1, w_acc);
                tc_fence_after_sync();
                step_ss(tmem_base + kColAcc2, true);
                step_ss(tmem_base + kColAcc2, false);
            };
            auto head = [&](int it) { head_m1(it); head_m2f(it); };
            head(0);
            for (int it = 0; it < my_tiles; ++it) {
                // M2X(it): needs xyz_feature in EA
                mbar_wait_timed<PROF>(ea_full, ea_use & 1, w_ea2);
                if (PROF) {
                    const long long now = clock64();
                    long long last = 0;
                    for (int w = 0; w < kEpiWarps; ++w) last = max(last, reinterpret_cast<volatile long long *>(smem + L.off_clk)[w]);
                    wake += (unsigned long long)(now - last);
                }
                tc_fence_after_sync();
                step_ts(tmem_base + kColAcc2, 0, false);
                step_ts(tmem_base + kColAcc2, 1, false);
                if (elect_one()) { mma_commit(&acc_full[1]); mma_commit(ea_empty); }
                __syncwarp();
                ++ea_use;
                if ((p.mode & 2) && it + 1 < my_tiles) head(it + 1);
                else if (it + 1 < my_tiles) head_m1(it + 1);
                // M3(it): needs merged in EA
                mbar_wait_timed<PROF>(ea_full, ea_use & 1, w_ea3);
                mbar_wait_timed<PROF>(&acc_empty[2], (uint32_t)(it & 1) ^ 1, w_acc);
                tc_fence_after_sync();
                step_ts(tmem_base + kColAcc3, 0, true);
                step_ts(tmem_base + kColAcc3, 1, false);
                if (elect_one()) { mma_commit(&acc_full[2]); mma_commit(ea_empty); }
                __syncwarp();
                ++ea_use;
                if (!(p.mode & 2) && it + 1 < my_tiles) head_m2f(it + 1);
            }
            if (PROF && lane == 0) {
                unsigned long long *o = p.prof + (size_t)blockIdx.x * 32;
                o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_a; o[2] = w_w; o[3] = w_ea2; o[4] = w_ea3; o[5] = w_acc;
                o[6] = (unsigned long long)my_tiles; o[7] = wake;
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        const float *bias = reinterpret_cast<const float *>(smem + L.off_bias);
        const int ew = warp - kFirstEpiWarp;
        const int q = ew & 3, half = ew >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t ea_use = 0;
        unsigned long long w_accf[3] = {0, 0, 0}, w_eae = 0, t_ea = 0, t_st = 0, t_ld12 = 0, t_ld3 = 0, t_stw = 0;
        // acc (128 columns) -> +b, ReLU -> bf16 hi / lo -> EA
        auto to_ea = [&](int a, int it, const float *b) {
            mbar_wait_timed<PROF>(&acc_full[a], (uint32_t)(it & 1), w_accf[a]);
            mbar_wait_timed<PROF>(ea_empty, (ea_use & 1) ^ 1, w_eae);
            const long long t0 = PROF ? clock64() : 0;
            tc_fence_after_sync();
            const uint32_t t_acc = lane_addr + (uint32_t)(a * 128);
            const uint32_t t_hi = lane_addr + kColEA, t_lo = t_hi + 64;
            auto convert = [&](const uint32_t (&v)[16], int c0) {
                uint32_t hi[8], lo[8];
                const float4 *b4 = reinterpret_cast<const float4 *>(b + c0);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 bb = b4[j4];
                    const float x0 = fmaxf(__uint_as_float(v[4 * j4 + 0]) + bb.x, 0.f);
                    const float x1 = fmaxf(__uint_as_float(v[4 * j4 + 1]) + bb.y, 0.f);
                    const float x2 = fmaxf(__uint_as_float(v[4 * j4 + 2]) + bb.z, 0.f);
                    const float x3 = fmaxf(__uint_as_float(v[4 * j4 + 3]) + bb.w, 0.f);
                    uint2 h2, l2;
                    split4(make_float4(x0, x1, x2, x3), h2, l2);
                    hi[2 * j4] = h2.x; hi[2 * j4 + 1] = h2.y;
                    lo[2 * j4] = l2.x; lo[2 * j4 + 1] = l2.y;
                }
                tmem_st8(t_hi + (c0 >> 1), hi);
                tmem_st8(t_lo + (c0 >> 1), lo);
            };
#pragma unroll 1
            for (int c0 = half * 16; c0 < kC; c0 += 64) {
                uint32_t va[16], vb[16];
                const long long tl0 = PROF ? clock64() : 0;
                tmem_ld16(t_acc + c0, va);
                tmem_ld16(t_acc + c0 + 32, vb);
                tmem_ld_wait();
                if (PROF) t_ld12 += (unsigned long long)(clock64() - tl0);
                convert(va, c0);
                convert(vb, c0 + 32);
            }
            const long long ts0 = PROF ? clock64() : 0;
            tmem_st_wait();
            if (PROF) t_stw += (unsigned long long)(clock64() - ts0);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (PROF) reinterpret_cast<volatile long long *>(smem + L.off_clk)[ew] = clock64();
                mbar_arrive(ea_full);
                mbar_arrive(&acc_empty[a]);
            }
            ++ea_use;
            if (PROF) t_ea += (unsigned long long)(clock64() - t0);
        };
        // acc3 -> +b1 -> H rows.  A thread owns one row of the tile (tcgen05.ld: lane = row) and 16 consecutive columns
        // per load, i.e. 64 contiguous bytes of its output row: four 16-byte stores, two per 32-byte sector, no shared
        // memory and no warp synchronisation (measured against the transpose-through-shared-memory variant of
        // linear_tc.cu, which spent 7.4 k cycles per tile here: see tools/prof_front.py)
        auto store_h = [&](int it) {
            const long long tile = first + (long long)it * stride;
            mbar_wait_timed<PROF>(&acc_full[2], (uint32_t)(it & 1), w_accf[2]);
            const long long t0 = PROF ? clock64() : 0;
            tc_fence_after_sync();
            const uint32_t t_acc = lane_addr + kColAcc3;
            const long long row0 = tile * BM + q * 32;
            const float *b = bias + 2 * kC;
            {
                const long long row = row0 + lane;
                float *yrow = p.h + row * p.ldh;
                const bool live = row < p.rows;
#pragma unroll 1
                for (int c0 = half * 32; c0 < kC; c0 += 64) {
                    uint32_t va[16], vb[16];
                    const long long tl0 = PROF ? clock64() : 0;
                    tmem_ld16(t_acc + c0, va);
                    tmem_ld16(t_acc + c0 + 16, vb);
                    tmem_ld_wait();
                    if (PROF) t_ld3 += (unsigned long long)(clock64() - tl0);
                    if (c0 + 64 >= kC) {      // the last columns of this warp are in registers: the accumulator may be reused
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[2]);
                    }
                    if (live) {
                        const float4 *b4 = reinterpret_cast<const float4 *>(b + c0);
                        float4 *y4 = reinterpret_cast<float4 *>(yrow + c0);
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 bb = b4[j4];
                            const uint32_t *v = j4 < 4 ? va + 4 * j4 : vb + 4 * (j4 - 4);
                            y4[j4] = make_float4(__uint_as_float(v[0]) + bb.x, __uint_as_float(v[1]) + bb.y,
                                                 __uint_as_float(v[2]) + bb.z, __uint_as_float(v[3]) + bb.w);
                        }
                    }
                }
            }
            if (PROF) t_st += (unsigned long long)(clock64() - t0);
        };
        if (my_tiles > 0) to_ea(0, 0, bias);                       // E1(0)
        for (int it = 0; it < my_tiles; ++it) {
            to_ea(1, it, bias + kC);                               // E2(it)
            if (it + 1 < my_tiles) to_ea(0, it + 1, bias);         // E1(it + 1)
            store_h(it);                                           // E3(it)
        }
        if (PROF && ew == 0 && lane == 0) {
            unsigned long long *o = p.prof + (size_t)blockIdx.x * 32;
            o[8] = w_accf[0]; o[9] = w_accf[1]; o[10] = w_accf[2]; o[11] = w_eae; o[12] = t_ea; o[13] = t_st;
            o[21] = t_ld12; o[22] = t_stw; o[23] = t_ld3;
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
int g_mode = 0;

}  // namespace

// stopwatch buffer for tools/prof_front.py: 32 u64 per CTA (device memory) or NULL to disable (never used by the product)
PN2_API void pn2_rcnn_front_set_profile(void *buf) { g_prof = static_cast<unsigned long long *>(buf); }
// tuning experiments only (tools/prof_front.py): bit1 = issue M2F before M3 (the first version's order)
PN2_API void pn2_rcnn_front_set_mode(int bits) { g_mode = bits; }

// The RCNN input chain in one launch (see the header of this file):
//   h[r] = W1f . relu(Wm . cat[relu(Wu2 . relu(Wpre . x[r, 0:5] + bpre) + bu2), x[r, off_f : off_f + 128]] + bm) + b1
// x (rows, ldx) pooled rows [5 extras | pad | 128 rpn features at column off_f], rows 16-byte aligned, off_f % 4 == 0;
// wpre (6, 128) f32: the five input-major weight rows of xyz_up layer 1, then its bias; w_up2 / w_merge / w_sa: the
// fused.pack_tc images of the (128 x 128), (128 x 256) and (128 x 128) weights (ntile 128; 2, 4, 2 K-blocks);
// h (rows, ldh) receives the PRE-activation per-point half of SA1's first layer.
PN2_API int pn2_rcnn_front_tc_f32(const float *x, int ldx, int off_f, const float *wpre, const void *w_up2,
                                  const float *b_up2, const void *w_merge, const float *b_merge, const void *w_sa,
                                  const float *b_sa, float *h, int ldh, long long rows, cudaStream_t stream) {
    if (!x || !wpre || !w_up2 || !b_up2 || !w_merge || !b_merge || !w_sa || !b_sa || !h || rows < 0 || ldh < kC ||
        off_f < 8 || ldx < off_f + kC) {
        pn2_set_last_error("pn2_rcnn_front_tc_f32: bad argument");
        return PN2_ERR_INVALID;
    }
    if ((ldx & 3) || (off_f & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (ldh & 3) || (reinterpret_cast<uintptr_t>(h) & 15) ||
        rows > 2147483647LL) {
        pn2_set_last_error("pn2_rcnn_front_tc_f32: rows of x and h must be 16-byte aligned");
        return PN2_ERR_UNSUPPORTED;
    }
    if (rows == 0) return PN2_OK;
    FrontParams p = {};
    p.x = x; p.ldx = ldx; p.off_f = off_f; p.rows = rows; p.tiles = (rows + BM - 1) / BM;
    p.wpre = wpre;
    p.w_up2 = static_cast<const uint8_t *>(w_up2); p.w_merge = static_cast<const uint8_t *>(w_merge);
    p.w_sa = static_cast<const uint8_t *>(w_sa);
    p.b_up2 = b_up2; p.b_merge = b_merge; p.b_sa = b_sa; p.h = h; p.ldh = ldh;
    const Smem L = make_layout();
    p.prof = g_prof;
    p.mode = g_mode;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(rcnn_front_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(rcnn_front_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_done = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(p.tiles < sms ? p.tiles : sms);
    if (p.prof) rcnn_front_tc_kernel<true><<<grid, kThreads, L.total + 1024, stream>>>(p);
    else rcnn_front_tc_kernel<false><<<grid, kThreads, L.total + 1024, stream>>>(p);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
