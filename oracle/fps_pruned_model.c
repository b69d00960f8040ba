/*
 * fps_pruned_model.c -- CPU model of the ALGORITHM of csrc/fps_cells.cu (furthest point sampling with exact spatial
 * pruning).  TEST INFRASTRUCTURE ONLY (tests/test_fps_pruning_model_cpu.py): it pins, without a GPU, the claim the kernel
 * rests on -- skipping every cell whose bounding-box lower bound, evaluated with the reference's own float expression, is
 * not below the cell's current maximum min-distance leaves the result of pointrcnn/pointnet2_lib/pointnet2/src/
 * sampling_gpu.cu:93-209 unchanged bit for bit, for ANY assignment of points to cells -- by comparing with orc_fps
 * (pn2_oracle.c, the literal restatement of that kernel).
 *
 * The model keeps the kernel's structure: cells of `cell` consecutive entries of a caller-given order, per cell the box of
 * its points and the exact maximum of their running distances, per round the test  lb(c, box) < cmax ; the winner of a
 * round is (max d2, min rank) with the reference's rank  bitrev(k mod bs) * ceil(N / bs) + k / bs  (header of csrc/fps.cu).
 * Build with -ffp-contract=off like the other oracle sources.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int orc_opt_n_threads(int work_size);

static inline float sqdist_m(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    t = fmaf(dz, dz, t);
    return t;
}

static uint32_t rank_of(int k, int log2bs, int cnt) {
    uint32_t tref = (uint32_t)k & ((1u << log2bs) - 1u), rev = 0;
    for (int i = 0; i < log2bs; ++i) rev |= ((tref >> i) & 1u) << (log2bs - 1 - i);
    return rev * (uint32_t)cnt + ((uint32_t)k >> log2bs);
}

/* xyz (n,3), order (n) a permutation, idx (m) out, temp (n) running distances (pre-filled, mutated like the reference's
 * scratch); returns the number of (round, cell) pairs that were touched (of (m-1) * ceil(n / cell)). */
long long orc_fps_pruned_model(const float *xyz, const int32_t *order, int n, int m, int cell, float *temp, int32_t *idx) {
    if (m <= 0 || n <= 0) return 0;
    const int bs = orc_opt_n_threads(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    const int cnt = (n + bs - 1) / bs;
    const int ncell = (n + cell - 1) / cell;
    float *lo = (float *)malloc(sizeof(float) * 3 * ncell), *hi = (float *)malloc(sizeof(float) * 3 * ncell);
    float *cmax = (float *)malloc(sizeof(float) * ncell);
    for (int c = 0; c < ncell; ++c) {
        for (int a = 0; a < 3; ++a) { lo[3 * c + a] = INFINITY; hi[3 * c + a] = -INFINITY; }
        cmax[c] = 0.f;
        for (int e = c * cell; e < n && e < (c + 1) * cell; ++e) {
            const int k = order[e];
            for (int a = 0; a < 3; ++a) {
                lo[3 * c + a] = fminf(lo[3 * c + a], xyz[3 * k + a]);
                hi[3 * c + a] = fmaxf(hi[3 * c + a], xyz[3 * k + a]);
            }
            cmax[c] = fmaxf(cmax[c], temp[k]);
        }
    }
    long long touched = 0;
    int old = 0;
    idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float cx = xyz[3 * old], cy = xyz[3 * old + 1], cz = xyz[3 * old + 2];
        for (int c = 0; c < ncell; ++c) {
            const float bx = fmaxf(fmaxf(lo[3 * c] + -cx, cx + -hi[3 * c]), 0.f);
            const float by = fmaxf(fmaxf(lo[3 * c + 1] + -cy, cy + -hi[3 * c + 1]), 0.f);
            const float bz = fmaxf(fmaxf(lo[3 * c + 2] + -cz, cz + -hi[3 * c + 2]), 0.f);
            if (!(sqdist_m(bx, by, bz) < cmax[c])) continue;          /* the cell cannot change: skipped */
            ++touched;
            float mx = 0.f;
            for (int e = c * cell; e < n && e < (c + 1) * cell; ++e) {
                const int k = order[e];
                const float d = sqdist_m(xyz[3 * k] + -cx, xyz[3 * k + 1] + -cy, xyz[3 * k + 2] + -cz);
                temp[k] = fminf(d, temp[k]);
                mx = fmaxf(mx, temp[k]);
            }
            cmax[c] = mx;
        }
        /* winner: maximum distance, then the reference's rank */
        int best = -1;
        float bd = -1.f;
        uint32_t br = 0;
        for (int k = 0; k < n; ++k) {
            const float d = temp[k];
            if (d > bd || (d == bd && rank_of(k, log2bs, cnt) < br)) { bd = d; best = k; br = rank_of(k, log2bs, cnt); }
        }
        old = best;
        idx[j] = best;
    }
    free(lo); free(hi); free(cmax);
    return touched;
}
