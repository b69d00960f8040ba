"""Index arithmetic of two kernels restated in Python and checked exhaustively without a GPU:
 * the register-blocked bitonic network of argsort_desc_kernel (csrc/glue.cu: sort_block_steps) sorts;
 * the inverted-rank table of csrc/fps_cells.cu is invertible (k_of_rinv(rinv_of(k)) == k) for every cloud size it takes,
   and real entries never collide with the padding value 0."""
import math

import numpy as np
import pytest


def _block_steps(keys, npow2, size, lo, nsub):
    e = 1 << nsub
    for blk in range(npow2 >> nsub):
        low, high = blk & ((1 << lo) - 1), blk >> lo
        base = (high << (lo + nsub)) | low
        up = (base & size) == 0
        idx = [base + (j << lo) for j in range(e)]
        r = [keys[i] for i in idx]
        for bit in range(nsub - 1, -1, -1):
            for j in range(e):
                if (j >> bit) & 1:
                    continue
                a, c = r[j], r[j | (1 << bit)]
                if (a > c) == up:
                    r[j], r[j | (1 << bit)] = c, a
        for i, v in zip(idx, r):
            keys[i] = v


def _network_sort(keys):
    npow2, s, size = len(keys), 1, 2
    while size <= npow2:
        hi, g0 = s, ((s - 1) & 3) + 1
        _block_steps(keys, npow2, size, hi - g0, g0)
        hi -= g0
        while hi > 0:
            _block_steps(keys, npow2, size, hi - 4, 4)
            hi -= 4
        size <<= 1
        s += 1


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 512, 2048])
def test_register_blocked_bitonic_network_sorts(n):
    rng = np.random.RandomState(n)
    keys = [int(v) for v in rng.randint(0, 1 << 20, size=n)]       # duplicates included
    want = sorted(keys)
    _network_sort(keys)
    assert keys == want


def _ref_block_size(n):
    v = 1 << int(math.log(n) / math.log(2.0))
    return max(1, min(v, 1024))


def _brev32(x):
    return int('{:032b}'.format(x)[::-1], 2)


@pytest.mark.parametrize("n", [1, 2, 3, 65, 1000, 1024, 1025, 2049, 4096, 9000, 16383, 16384])
def test_inverted_rank_table_is_invertible(n):
    bs = _ref_block_size(n)
    log2bs = bs.bit_length() - 1
    cnt = (n + bs - 1) // bs
    seen = set()
    for k in range(n):
        tref = k & ((1 << log2bs) - 1)
        rev = (_brev32(tref) >> (32 - log2bs)) if log2bs else 0
        rank = rev * cnt + (k >> log2bs)
        rinv = ~rank & 0xFFFF
        assert rinv != 0 and rank < 0xFFFF            # 0 is the padding value
        assert rinv not in seen                       # ranks are unique: the tie-break is total
        seen.add(rinv)
        rank2 = ~rinv & 0xFFFF
        if log2bs == 0:
            k2 = rank2
        else:
            rev2, kd = divmod(rank2, cnt)
            k2 = (kd << log2bs) | (_brev32(rev2) >> (32 - log2bs))
        assert k2 == k
