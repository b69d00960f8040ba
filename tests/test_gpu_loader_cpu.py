"""Host half of the GPU data path (datasets/gpu_loader.py): draw_selection replays the np.random calls of
KittiRCNNDataset._sample_indices (pointrcnn/lib/datasets/kitti_rcnn_dataset.py:291-320) on COUNTS; resolved
against the ordered near / far lists it must pick exactly the rows `pts_rect[choice]` picks.  No GPU needed."""
import numpy as np
import pytest

from conftest import load


def _resolve(sel, near_list, far_list):
    gl = load("datasets.gpu_loader")
    out = np.empty(len(sel), np.int64)
    for k, s in enumerate(sel):
        s = int(s)
        if s < 0:
            out[k] = -s - 1
        elif s < gl.FAR_BASE:
            out[k] = near_list[s]
        else:
            out[k] = far_list[s - gl.FAR_BASE]
    return out


@pytest.mark.parametrize("n_valid,far_frac,npoints,faraway", [
    (30000, 0.30, 16384, 4000),      # more far points than the cap: far subsampled, near subsampled
    (30000, 0.05, 16384, 4000),      # few far points: all kept
    (20000, 0.70, 16384, 4000),      # near < needed: near drawn WITH replacement
    (9000, 0.20, 16384, 4000),       # fewer valid points than npoints: pad by duplication (replace=False)
    (5000, 0.20, 16384, 4000),       # ... fewer than the missing count: padding drawn with replacement
    (16384, 0.10, 16384, 4000),      # exactly npoints: only the shuffle
    (40000, 0.0, 16384, 4000),       # no far point at all
])
def test_draw_selection_replays_sample_indices(n_valid, far_frac, npoints, faraway):
    load("config").use_default_yaml("rcnn")
    ds_mod, gl = load("datasets.kitti_rcnn_dataset"), load("datasets.gpu_loader")
    rng = np.random.RandomState(n_valid)
    pts_rect = rng.uniform(-30, 30, (n_valid, 3)).astype(np.float32)
    far = rng.random_sample(n_valid) < far_frac
    pts_rect[:, 2] = np.where(far, rng.uniform(40, 70, n_valid), rng.uniform(0, 39.9, n_valid)).astype(np.float32)
    ds = ds_mod.KittiRCNNDataset.__new__(ds_mod.KittiRCNNDataset)
    ds.npoints, ds.npoints_faraway, ds.with_replace = npoints, faraway, False
    near_list = np.where(pts_rect[:, 2] < 40.0)[0]
    far_list = np.where(~(pts_rect[:, 2] < 40.0))[0]
    for seed in (666, 1):
        np.random.seed(seed)
        want = ds._sample_indices(pts_rect)
        state_ref = np.random.get_state()[1].copy()
        np.random.seed(seed)
        sel = gl.draw_selection(n_valid, len(near_list), len(far_list), npoints, faraway, False)
        state_new = np.random.get_state()[1].copy()
        assert sel.dtype == np.int32 and sel.shape == (npoints,)
        assert np.array_equal(_resolve(sel, near_list, far_list), np.asarray(want, np.int64))
        assert np.array_equal(state_ref, state_new)          # the generator is left in the same state: scene order holds


@pytest.mark.parametrize("n_valid,n_far,npoints,faraway,with_replace", [
    (30000, 9000, 16384, 4000, False), (30000, 1500, 16384, 4000, False), (20000, 14000, 16384, 4000, False),
    (9000, 1800, 16384, 4000, False), (5000, 1000, 16384, 4000, False), (16384, 1600, 16384, 4000, False),
    (40000, 0, 16384, 4000, False), (30000, 9000, 16384, 4000, True), (100000, 30000, 16384, 4000, False),
    (17, 3, 64, 8, False), (1, 0, 16, 4, False)])
def test_native_mt19937_draws_equal_numpy(n_valid, n_far, npoints, faraway, with_replace):
    """csrc/mt_select.cu (host code in libpn2_b200.so) against numpy's legacy RandomState: same encoded selection,
    same generator state afterwards -- seeded per scene and continuing the global np.random stream."""
    gl = load("datasets.gpu_loader")
    n_near = n_valid - n_far
    for seed in (666, 0, 2 ** 32 - 1):
        np.random.seed(seed)
        np.random.random_sample(5)                      # a stream that is already under way (pos != 624)
        scratch = np.empty((max(n_valid, npoints) + npoints,), np.int32)
        for rep in range(3):                            # consecutive scenes on one stream
            st = gl.MTState.from_numpy_global()
            want = gl.draw_selection(n_valid, n_near, n_far, npoints, faraway, with_replace)
            after = np.random.get_state()
            got = np.empty((npoints,), np.int32)
            gl.draw_selection_native(st, n_valid, n_near, n_far, npoints, faraway, with_replace, got, scratch)
            assert np.array_equal(got, want)
            assert np.array_equal(st.key, after[1]) and int(st.pos.value) == after[2]
        rs = np.random.RandomState(seed)
        want = gl.draw_selection(n_valid, n_near, n_far, npoints, faraway, with_replace, rng=rs)
        st = gl.MTState.seeded(seed)
        gl.draw_selection_native(st, n_valid, n_near, n_far, npoints, faraway, with_replace, got, scratch)
        assert np.array_equal(got, want)
        assert np.array_equal(st.key, rs.get_state()[1]) and int(st.pos.value) == rs.get_state()[2]
    # to_numpy_global hands the stream back
    np.random.seed(3)
    st = gl.MTState.from_numpy_global()
    gl.draw_selection_native(st, n_valid, n_near, n_far, npoints, faraway, with_replace, got, scratch)
    st.to_numpy_global()
    a = np.random.random_sample(4)
    np.random.seed(3)
    gl.draw_selection(n_valid, n_near, n_far, npoints, faraway, with_replace)
    assert np.array_equal(a, np.random.random_sample(4))


def test_native_mt19937_draws_fuzz_small_cases():
    """Random small (n_valid, n_far, npoints, npoints_faraway, with_replace) against numpy, INCLUDING the corners the
    reference configuration never reaches (npoints_faraway >= npoints, so no near point is needed: numpy's
    choice(n, 0, replace=False) still draws the whole permutation and the stream must advance the same way) and the
    cases where numpy raises (nothing to draw from): there the native call must fail too, never invent a selection."""
    gl = load("datasets.gpu_loader")
    rs = np.random.RandomState(0)
    n_ok = n_err = 0
    for _ in range(4000):
        n_valid = int(rs.randint(0, 120))
        n_far = int(rs.randint(0, n_valid + 1))
        npoints = int(rs.randint(1, 100))
        faraway = int(rs.randint(0, npoints + 3))
        with_replace = bool(rs.randint(2))
        seed = int(rs.randint(0, 2 ** 31))
        case = (n_valid, n_far, npoints, faraway, with_replace, seed)
        ref = np.random.RandomState(seed)
        try:
            want = gl.draw_selection(n_valid, n_valid - n_far, n_far, npoints, faraway, with_replace, rng=ref)
        except ValueError:
            want = None
        st = gl.MTState.seeded(seed)
        got = np.empty((npoints,), np.int32)
        scratch = np.empty((max(n_valid, npoints) + npoints,), np.int32)
        if want is None:
            with pytest.raises(Exception):
                gl.draw_selection_native(st, n_valid, n_valid - n_far, n_far, npoints, faraway, with_replace, got, scratch)
            n_err += 1
            continue
        gl.draw_selection_native(st, n_valid, n_valid - n_far, n_far, npoints, faraway, with_replace, got, scratch)
        assert np.array_equal(got, want), case
        assert np.array_equal(st.key, ref.get_state()[1]) and int(st.pos.value) == ref.get_state()[2], case
        n_ok += 1
    assert n_ok > 3000 and n_err > 10          # both kinds of case were actually exercised
