// mt_select.cu -- HOST code: the np.random draws of KittiRCNNDataset._sample_indices
// (pointrcnn/lib/datasets/kitti_rcnn_dataset.py:291-320) replayed on an MT19937 state, bit for bit what numpy's
// legacy RandomState does (numpy/random/mtrand.pyx: choice -> permutation -> shuffle -> random_interval;
// randint -> masked rejection), without the interpreter.
//
// Why native: which points a scene keeps is DEFINED by that generator stream, so the draws cannot move to the
// GPU, and in Python they cost 1.8-2.3 ms per scene (a 50 k-element permutation through numpy's generic
// byte-wise swap loop + a 16384-element shuffle) -- more than the whole GPU forward pass of the scene (0.57 ms).
// The generator state is handed over explicitly (key[624], pos = np.random.get_state()[1:3]), so the global
// np.random stream of the reference stays intact across calls, and because ctypes releases the GIL scenes with
// their own seed are drawn on several host threads (datasets/gpu_loader.py).
//
// Checked against numpy itself (tests/test_gpu_loader_cpu.py): selections and final generator state equal.
#include "common.cuh"
#include <stdint.h>
#include <string.h>

namespace {

struct MT {
    uint32_t *key;   // 624 words, caller-owned
    int pos;
};

inline void mt_gen(MT &s) {
    const uint32_t N = 624, M = 397, MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t *mt = s.key;
    uint32_t y;
    uint32_t kk = 0;
    for (; kk < N - M; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < N - 1; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    s.pos = 0;
}

inline uint32_t mt_next32(MT &s) {
    if (s.pos == 624) mt_gen(s);
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}


// numpy/random/src/distributions/distributions.c: random_interval (legacy masked rejection): the smallest bit mask
// covering max, then draws until one falls into [0, max].  (All populations here are < 2^31: the 32-bit branch.)
inline uint32_t mask_of(uint32_t max) {
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    return mask;
}
inline uint32_t interval(MT &s, uint32_t max) {
    if (max == 0) return 0;
    const uint32_t mask = mask_of(max);
    uint32_t value;
    while ((value = (mt_next32(s) & mask)) > max) {}
    return value;
}

// RandomState.shuffle on a 1-D array: for i = n-1 .. 1: j = random_interval(i); swap(x[i], x[j]).
// The j sequence depends on the generator stream and on i only, never on x.  So the draws run ahead of the swaps in
// blocks: a branch-free rejection loop (the accept bit advances the output slot and the bound; no data-dependent
// branch to mispredict at the ~25 % rejection rate) fills a block of accepted j's, then the swaps of the block run
// with all their addresses known, so the cache misses of the random side overlap.  Same draws, same order, the
// generator stops right after the last accepted value -- bit for bit numpy's result and final state, 2.1x faster
// than the draw-then-swap loop on a 50 k-point scene (0.96 -> 0.45 ms for 35 k + 15 k permutations and the 16384 shuffle).
inline void shuffle(MT &s, int32_t *__restrict__ x, long long n) {
    if (n < 2) return;
    constexpr int kBlock = 1024;
    uint32_t js[kBlock];
    uint32_t *__restrict__ key = s.key;
    int pos = s.pos;
    long long i = n - 1;
    while (i > 0) {
        const int m = i < kBlock ? (int)i : kBlock;      // swaps of this block: positions i, i-1, ..., i-m+1
        int cnt = 0;
        uint32_t bound = (uint32_t)i;                    // random_interval(bound): mask = smallest 2^k - 1 >= bound
        while (cnt < m) {
            if (pos == 624) { s.pos = pos; mt_gen(s); pos = 0; }
            uint32_t y = key[pos++];
            y ^= (y >> 11);
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= (y >> 18);
            const uint32_t v = y & (0xffffffffu >> __builtin_clz(bound));
            js[cnt] = v;
            const uint32_t accept = v <= bound;
            cnt += (int)accept;
            bound -= accept;                             // bound >= 1 while cnt < m
        }
        for (int k = 0; k < m; ++k) {
            const long long a = i - k;
            const uint32_t j = js[k];
            const int32_t t = x[a]; x[a] = x[j]; x[j] = t;
        }
        i -= m;
    }
    s.pos = pos;
}

// RandomState.choice(n, size, replace=False) = permutation(n)[:size]: the whole permutation is drawn
inline void choice_no_replace(MT &s, long long n, int32_t *scratch) {
    for (long long i = 0; i < n; ++i) scratch[i] = (int32_t)i;
    shuffle(s, scratch, n);
}

// RandomState.choice(n, size, replace=True) = randint(0, n, size): legacy masked rejection on [0, n-1]
inline void choice_replace(MT &s, long long n, long long size, int32_t *out) {
    const uint32_t rng = (uint32_t)(n - 1);
    for (long long i = 0; i < size; ++i) out[i] = rng == 0 ? 0 : (int32_t)interval(s, rng);
}

}  // namespace

// RandomState(seed) / np.random.seed(seed) for an integer seed: init_genrand, pos = 624.
PN2_API void pn2_mt_seed(uint32_t seed, uint32_t *key, int32_t *pos) {
    key[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) key[i] = 1812433253u * (key[i - 1] ^ (key[i - 1] >> 30)) + i;
    *pos = 624;
}

// The draws of _sample_indices for one scene, from the counts of pn2_scene_filter_f32, encoded for
// pn2_scene_gather_f32 ([0, 2^30) near_list index, [2^30, 2^31) far_list index, negative = -(valid index) - 1).
// key (624) / pos: MT19937 state, updated in place.  scratch: max(n_valid, npoints) + npoints int32.
// Returns PN2_OK, or PN2_ERR_INVALID for an empty population that would have to be sampled (numpy raises there).
PN2_API int pn2_mt_draw_selection(uint32_t *key, int32_t *pos, int n_valid, int n_near, int n_far, int npoints,
                                  int npoints_faraway, int with_replace, int32_t *sel, int32_t *scratch) {
    if (!key || !pos || !sel || !scratch || n_valid < 0 || n_near < 0 || n_far < 0 || npoints <= 0 || n_near + n_far != n_valid ||
        n_valid >= (1 << 30)) {
        pn2_set_last_error("pn2_mt_draw_selection: bad argument");
        return PN2_ERR_INVALID;
    }
    MT s = {key, *pos};
    const long long pop = n_valid > npoints ? n_valid : npoints;
    int32_t *choice = scratch + pop;            // npoints entries: the selection before the final shuffle
    const int32_t FAR_BASE = 1 << 30;
    long long len = 0;
    if (npoints < n_valid) {
        // far points: all of them, or npoints_faraway of a permutation
        long long n_far_sel = n_far;
        const bool far_sub = n_far > npoints_faraway;
        if (far_sub) {
            choice_no_replace(s, n_far, scratch);
            n_far_sel = npoints_faraway;
        }
        const long long need = npoints - n_far_sel;
        if (need < 0) {
            pn2_set_last_error("pn2_mt_draw_selection: npoints_faraway exceeds npoints");
            return PN2_ERR_INVALID;
        }
        // the far selection is parked at the END of `choice` while the near draw uses the scratch
        for (long long i = 0; i < n_far_sel; ++i) choice[need + i] = (far_sub ? scratch[i] : (int32_t)i) + FAR_BASE;
        if (need > 0 && n_near == 0) {
            pn2_set_last_error("pn2_mt_draw_selection: no near point to draw from");
            return PN2_ERR_INVALID;
        }
        if (n_near < need || with_replace) {
            choice_replace(s, n_near, need, choice);
        } else {
            // also for need == 0: choice(n, 0, replace=False) is permutation(n)[:0] -- the generator still advances
            choice_no_replace(s, n_near, scratch);
            memcpy(choice, scratch, (size_t)need * sizeof(int32_t));
        }
        len = npoints;      // near (need) followed by far (n_far_sel): np.concatenate((near, far))
    } else {
        for (long long i = 0; i < n_valid; ++i) choice[i] = (int32_t)(-i - 1);
        len = n_valid;
        if (npoints > n_valid) {
            const long long missing = npoints - n_valid;
            if (n_valid == 0) {
                pn2_set_last_error("pn2_mt_draw_selection: empty scene");
                return PN2_ERR_INVALID;
            }
            if (n_valid < missing) {
                choice_replace(s, n_valid, missing, scratch);
            } else {
                choice_no_replace(s, n_valid, scratch);
            }
            for (long long i = 0; i < missing; ++i) choice[n_valid + i] = -scratch[i] - 1;
            len = npoints;
        }
    }
    shuffle(s, choice, len);                    // np.random.shuffle(choice)
    memcpy(sel, choice, (size_t)len * sizeof(int32_t));
    *pos = s.pos;
    return PN2_OK;
}
