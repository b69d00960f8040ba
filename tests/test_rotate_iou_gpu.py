"""rotate_iou_gpu_eval (evaluate/rotate_iou.py) on the sm_100a kernel against golden matrices produced
on a B200 by the UNMODIFIED reference file under numba.cuda (tools/make_goldens.py ->
tests/golden/rotate_iou_numba.npz).  BASELINE.json config 2: 1000 x 1000, bit-exact."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rotate_iou_numba.npz")


def rotate_iou_inputs(seed, n):       # same generator as tools/make_goldens.py
    rng = np.random.RandomState(seed)
    c = rng.uniform(-5, 5, size=(n, 2))
    d = rng.uniform(1, 4, size=(n, 2))
    a = rng.uniform(-np.pi, np.pi, size=(n, 1))
    return np.concatenate([c, d, a], 1).astype(np.float32)


def mismatch(got, ref):
    bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
    return int(bad.sum()), (float(np.nanmax(np.abs(got[bad] - ref[bad]))) if bad.any() else 0.0)


@pytest.mark.parametrize("crit", [-1, 0, 1, 2])
def test_small_and_adversarial_bit_exact(cuda, crit):
    g = np.load(GOLD)
    f = load("rotate_iou").rotate_iou_gpu_eval
    a, b = rotate_iou_inputs(0, 1000)[:160], rotate_iou_inputs(1, 1000)[:130]
    got = f(a, b, crit)
    assert got.dtype == np.float32 and got.shape == (160, 130)
    assert mismatch(got, g["small_c%d" % crit]) == (0, 0.0)
    # adversarial pairs (identical boxes, shared edges, zero area, contained, axis-aligned): bit-exact
    # wherever the reference is defined.  It keeps the polygon in a local array of 8 points
    # (rotate_iou.py:233); pairs with more vertices (corners inside + edge crossings, duplicates
    # counted) write out of bounds there, so its value is whatever lies next on the numba stack frame.
    from oracle import oracle as orc
    adv = g["adv"]
    _, npts = orc.rotate_iou_eval(adv, adv, crit, return_npts=True)
    got, ref = f(adv, adv, crit), g["adv_c%d" % crit]
    defined = npts <= 8
    assert defined.sum() >= 200 and (~defined).sum() > 0
    assert mismatch(np.where(defined, got, 0), np.where(defined, ref, 0)) == (0, 0.0)
    assert np.isfinite(got[~defined]).all()      # (duplicate vertices make the fan area exceed the box: not an IoU)


@pytest.mark.parametrize("crit", [-1, 0, 1, 2])
def test_config2_1000x1000_bit_exact(cuda, crit):
    g = np.load(GOLD)
    f = load("rotate_iou").rotate_iou_gpu_eval
    big = f(rotate_iou_inputs(0, 1000), rotate_iou_inputs(1, 1000), crit)
    assert mismatch(big[::7, ::11], g["big_diag_c%d" % crit]) == (0, 0.0)
    assert np.array_equal(np.frombuffer(hashlib.sha256(big.tobytes()).digest(), np.uint8), g["big_sha_c%d" % crit])


def test_contract_dtype_and_empty(cuda):
    f = load("rotate_iou").rotate_iou_gpu_eval
    a = rotate_iou_inputs(3, 7).astype(np.float64)
    out = f(a, a)
    # float32 whatever the input dtype: rotate_iou.py:312 rebinds `boxes` to its float32 cast before :329 reads boxes.dtype
    assert out.dtype == np.float32 and out.shape == (7, 7)
    assert f(a[:0], a).shape == (0, 7) and f(a[:0], a).dtype == np.float32   # early return keeps float32 (:315-316)
    assert f(a, a[:0]).shape == (7, 0)
    # properties: symmetric IoU under swapping the roles, inter <= min area
    inter = f(a, a, 2)
    areas = (a[:, 2] * a[:, 3]).astype(np.float32)
    assert (inter <= np.minimum(areas[:, None], areas[None, :]) * (1 + 1e-5) + 1e-6).all()
    np.testing.assert_allclose(f(a, a, -1), f(a, a, -1).T, rtol=1e-5, atol=1e-6)
