// group_compact.cu -- the UNIQUE rows of ball-query groups, for the duplicate-skipping SA kernels.
//
// ball_query pads a group that found cnt < nsample neighbours with copies of its first hit
// (pointnet2/src/ball_query_gpu.cu:35-39: the first hit fills every slot, later hits overwrite slots 1..cnt-1), and the
// reference then pushes all nsample rows through the shared MLP and max-pools them (pointnet2_modules.py:38-44).  The
// padded rows are bit-for-bit copies of row 0, so they cannot change the maximum: only the first cnt rows of a group
// carry information.  On the benchmark clouds 44-92 % of all grouped rows are such copies (tools/bq_fill_stats.py).
// Since the hits of a group are distinct point indices, cnt = 1 + #{k >= 1 : idx[k] != idx[0]} and the unique rows are
// exactly slots 0..cnt-1.  (A group without any hit keeps its zero-initialised row: cnt = 1, neighbour 0, like the
// reference, which gathers point 0 sixty-four times.)
//   pn2_group_unique_count_i32: cnt (G) from idx (G, ns), rounded up to a multiple of `align`
//   pn2_group_compact_i32     : with the exclusive prefix sum of cnt, the compact row list
//                               cmap[u] = group (centre) of unique row u, jmap[u] = its neighbour index.
// align (1, 2, 4, 8 or 16, ns % align == 0): a group's rows are topped up to a multiple of `align` with further copies
// of its row 0 -- still no influence on the maximum -- so that every aligned run of `align` list rows belongs to ONE
// group.  The transposed SA kernel pools such a list eight columns at a time without looking at per-column centre ids
// (csrc/sa_fused_t_tc.cu); the price is (align - 1) / 2 extra rows per group on average.
// Both enqueue on the stream and never synchronise; the total U = sum(cnt) stays on the device (the SA kernel reads it).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) unique_count_kernel(const int32_t *__restrict__ idx, long long g, int ns, int align,
                                                          int32_t *__restrict__ cnt) {
    const long long grp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (grp >= g) return;
    const int32_t *row = idx + grp * ns;
    const int32_t first = __ldg(row);
    int c = 0;
    for (int k = lane; k < ns; k += 32) c += (k == 0 || __ldg(row + k) != first) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[grp] = (c + align - 1) / align * align;
}

__global__ void __launch_bounds__(256) compact_kernel(const int32_t *__restrict__ idx, long long g, int ns,
                                                     const int32_t *__restrict__ cnt, const long long *__restrict__ offs,
                                                     int32_t *__restrict__ cmap, int32_t *__restrict__ jmap) {
    const long long grp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (grp >= g) return;
    const long long o = offs[grp];
    const int32_t *row = idx + grp * ns;
    const int32_t first = __ldg(row);
    int base = 0;
    for (int k0 = 0; k0 < ns; k0 += 32) {         // ordered: slot 0 and every slot that differs from it (for ball_query
        const int k = k0 + lane;                  // output these are the leading cnt slots; any idx tensor is handled)
        const int32_t v = k < ns ? __ldg(row + k) : first;
        const bool keep = k < ns && (k == 0 || v != first);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const long long pos = o + base + __popc(bal & ((1u << lane) - 1u));
            cmap[pos] = (int32_t)grp;
            jmap[pos] = v;
        }
        base += __popc(bal);
    }
    // rows up to the (aligned) count: further copies of row 0
    for (int k = base + lane; k < __ldg(cnt + grp); k += 32) {
        cmap[o + k] = (int32_t)grp;
        jmap[o + k] = first;
    }
}

// ---- the whole compaction in two launches (count + block sums | offsets + lists) instead of count, a two-launch
//      device scan, a subtraction and the list kernel: a CTA takes kBlockGroups consecutive groups, the exclusive offset of
//      its first group is the sum of the block totals before it (at most a few hundred values) ----
constexpr int kBlockGroups = 256;       // 8 warps x 32 groups

__global__ void __launch_bounds__(256) unique_count_blocks_kernel(const int32_t *__restrict__ idx, long long g, int ns, int align,
                                                                 int32_t *__restrict__ cnt, int32_t *__restrict__ block_sum) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long g0 = (long long)blockIdx.x * kBlockGroups + warp * 32;
    int mine = 0;                        // lane i keeps the count of group g0 + i
    for (int i = 0; i < 32; ++i) {
        const long long grp = g0 + i;
        if (grp >= g) break;             // warp-uniform
        const int32_t *row = idx + grp * ns;
        const int32_t first = __ldg(row);
        int c = 0;
        for (int k = lane; k < ns; k += 32) c += (k == 0 || __ldg(row + k) != first) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        c = (c + align - 1) / align * align;
        if (lane == i) mine = c;
    }
    if (g0 + lane < g) cnt[g0 + lane] = mine;
    int s = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        block_sum[blockIdx.x] = t;
    }
}

// the same counts from the ball query's own hit counts (pn2_ball_query_culled_fill_f32): unique rows = max(hits, 1) -- a group
// without a neighbour keeps its single zero row -- rounded up to the alignment; reads 4 bytes per group instead of the lists
__global__ void __launch_bounds__(256) counts_from_hits_kernel(const int32_t *__restrict__ hits, long long g, int align,
                                                              int32_t *__restrict__ cnt, int32_t *__restrict__ block_sum) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long grp = (long long)blockIdx.x * kBlockGroups + threadIdx.x;
    int c = 0;
    if (grp < g) {
        c = (max(__ldg(hits + grp), 1) + align - 1) / align * align;
        cnt[grp] = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        block_sum[blockIdx.x] = t;
    }
}

// HITS: the lists come from ball_query with its hit counts -- the unique rows of a group are its first `hits` slots, so a
// list row is a plain copy (no ballot compaction), and four groups are in flight per warp
template <bool HITS>
__global__ void __launch_bounds__(256) compact_blocks_kernel(const int32_t *__restrict__ idx, long long g, int ns,
                                                            const int32_t *__restrict__ cnt, const int32_t *__restrict__ block_sum,
                                                            const int32_t *__restrict__ hits,
                                                            int32_t *__restrict__ cmap, int32_t *__restrict__ jmap,
                                                            long long *__restrict__ total) {
    __shared__ long long red[8];
    __shared__ long long offs_s[kBlockGroups];
    __shared__ int wtot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // exclusive offset of this CTA's first group
    long long base = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += 256) base += __ldg(block_sum + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    if (lane == 0) red[warp] = base;
    // exclusive scan of the CTA's counts (thread t <-> group blockIdx.x * 256 + t)
    const long long gt = (long long)blockIdx.x * kBlockGroups + threadIdx.x;
    const int c = gt < g ? __ldg(cnt + gt) : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    base = 0;
    for (int w = 0; w < 8; ++w) base += red[w];
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wtot[w];
    offs_s[threadIdx.x] = base + wbase + inc - c;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) *total = base + wbase + inc;      // sum of all counts
    __syncthreads();
    if constexpr (HITS) {
        const long long gw = (long long)blockIdx.x * kBlockGroups + warp * 32;      // first group of this warp
        const int h = gw + lane < g ? max(__ldg(hits + gw + lane), 1) : 0;           // lane i: unique rows of group i
        constexpr int kG = 4;
        for (int i0 = 0; i0 < 32 && gw + i0 < g; i0 += kG) {
            int32_t v[kG][2], first[kG];
            int cg[kG], hg[kG];
            long long o[kG];
#pragma unroll
            for (int u = 0; u < kG; ++u) {
                const int i = i0 + u;
                const bool live = gw + i < g;
                cg[u] = live ? __shfl_sync(0xffffffffu, c, i) : 0;
                hg[u] = __shfl_sync(0xffffffffu, h, i);
                o[u] = offs_s[warp * 32 + i];
                const int32_t *row = idx + (gw + (live ? i : 0)) * ns;
                first[u] = __ldg(row);
                v[u][0] = lane < hg[u] ? __ldg(row + lane) : first[u];
                v[u][1] = lane + 32 < hg[u] ? __ldg(row + lane + 32) : first[u];
            }
#pragma unroll
            for (int u = 0; u < kG; ++u) {
                const int32_t grp = (int32_t)(gw + i0 + u);
                if (lane < cg[u]) { cmap[o[u] + lane] = grp; jmap[o[u] + lane] = v[u][0]; }
                if (lane + 32 < cg[u]) { cmap[o[u] + lane + 32] = grp; jmap[o[u] + lane + 32] = v[u][1]; }
                for (int k = lane + 64; k < cg[u]; k += 32) {      // nsample 128
                    cmap[o[u] + k] = grp;
                    jmap[o[u] + k] = k < hg[u] ? __ldg(idx + (long long)grp * ns + k) : first[u];
                }
            }
        }
    } else {
    // the lists of the warp's 32 groups (compact_kernel, one group at a time)
    for (int i = 0; i < 32; ++i) {
        const long long grp = (long long)blockIdx.x * kBlockGroups + warp * 32 + i;
        if (grp >= g) break;
        const long long o = offs_s[warp * 32 + i];
        const int32_t *row = idx + grp * ns;
        const int32_t first = __ldg(row);
        int nb = 0;
        for (int k0 = 0; k0 < ns; k0 += 32) {
            const int k = k0 + lane;
            const int32_t v = k < ns ? __ldg(row + k) : first;
            const bool keep = k < ns && (k == 0 || v != first);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const long long pos = o + nb + __popc(bal & ((1u << lane) - 1u));
                cmap[pos] = (int32_t)grp;
                jmap[pos] = v;
            }
            nb += __popc(bal);
        }
        const int cg = __shfl_sync(0xffffffffu, c, i);            // this group's (aligned) count lives in lane i
        for (int k = nb + lane; k < cg; k += 32) {
            cmap[o + k] = (int32_t)grp;
            jmap[o + k] = first;
        }
    }
    }
}

}  // namespace

// pn2_group_unique_count_i32 + exclusive scan + pn2_group_compact_i32 in two launches.  cnt (G) int32 and block_sum
// (ceil(G / 256)) int32 are scratch; cmap / jmap: capacity >= G * ns; *total (device, int64) = number of list rows.
// hits (G) int32 or NULL: the ball query's hit counts; with them the count pass reads 4 bytes per group, not the lists.
PN2_API int pn2_group_compact_lists_i32(const int32_t *idx, long long g, int ns, int align, const int32_t *hits, int32_t *cnt,
                                        int32_t *block_sum, int32_t *cmap, int32_t *jmap, long long *total, cudaStream_t stream) {
    if (g <= 0 || ns <= 0 || !idx || !cnt || !block_sum || !cmap || !jmap || !total || g > 2147483647LL || align < 1 ||
        align > 16 || (align & (align - 1)) || ns % align) {
        pn2_set_last_error("pn2_group_compact_lists_i32: bad argument");
        return PN2_ERR_INVALID;
    }
    const unsigned blocks = (unsigned)((g + kBlockGroups - 1) / kBlockGroups);
    if (hits) counts_from_hits_kernel<<<blocks, 256, 0, stream>>>(hits, g, align, cnt, block_sum);
    else unique_count_blocks_kernel<<<blocks, 256, 0, stream>>>(idx, g, ns, align, cnt, block_sum);
    PN2_CHECK_LAUNCH();
    if (hits) compact_blocks_kernel<true><<<blocks, 256, 0, stream>>>(idx, g, ns, cnt, block_sum, hits, cmap, jmap, total);
    else compact_blocks_kernel<false><<<blocks, 256, 0, stream>>>(idx, g, ns, cnt, block_sum, nullptr, cmap, jmap, total);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

PN2_API int pn2_group_unique_count_i32(const int32_t *idx, long long g, int ns, int align, int32_t *cnt,
                                       cudaStream_t stream) {
    if (g < 0 || ns <= 0 || (g > 0 && (!idx || !cnt)) || align < 1 || align > 16 || (align & (align - 1)) || ns % align) {
        pn2_set_last_error("pn2_group_unique_count_i32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (g == 0) return PN2_OK;
    unique_count_kernel<<<(unsigned)((g + 7) / 8), 256, 0, stream>>>(idx, g, ns, align, cnt);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}

// cnt (G): pn2_group_unique_count_i32's output (with its align); offs (G) int64: EXCLUSIVE prefix sum of cnt;
// cmap / jmap: capacity >= sum(cnt)
PN2_API int pn2_group_compact_i32(const int32_t *idx, long long g, int ns, const int32_t *cnt, const long long *offs,
                                  int32_t *cmap, int32_t *jmap, cudaStream_t stream) {
    if (g < 0 || ns <= 0 || (g > 0 && (!idx || !cnt || !offs || !cmap || !jmap)) || g > 2147483647LL) {
        pn2_set_last_error("pn2_group_compact_i32: bad argument");
        return PN2_ERR_INVALID;
    }
    if (g == 0) return PN2_OK;
    compact_kernel<<<(unsigned)((g + 7) / 8), 256, 0, stream>>>(idx, g, ns, cnt, offs, cmap, jmap);
    PN2_CHECK_LAUNCH();
    return PN2_OK;
}
