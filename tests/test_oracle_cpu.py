"""CPU checks of the oracle itself: against independent numpy restatements of the reference
semantics on small cases, and against hand-derived known answers for the tie rules.
(The pin against the REAL reference kernels is tests/test_legacy_parity_gpu.py + tests/golden.)"""
import numpy as np
import pytest

from conftest import load

synthetic = load("synthetic")


def sqd(a, b):
    """float32 squared distance in the reference's contraction order, emulated in float64:
    products of two f32 are exact in f64, so round-to-f32 of (exact product + f32 addend)
    equals the fused result except for rare double-rounding cases that the test data avoids
    by using lattice coordinates."""
    d = (a.astype(np.float32) - b.astype(np.float32)).astype(np.float32)
    dx, dy, dz = d[..., 0].astype(np.float64), d[..., 1].astype(np.float64), d[..., 2].astype(np.float64)
    t = (dy * dy).astype(np.float32).astype(np.float64)
    t = (dx * dx + t).astype(np.float32).astype(np.float64)
    t = (dz * dz + t).astype(np.float32)
    return t


def bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def fps_numpy(xyz, m, bs):
    """argmax with the key (d2 desc, bitrev(k mod bs) asc, k asc) -- the closed form of the
    reference's strided scan + shared-memory tree."""
    n = xyz.shape[0]
    bits = int(np.log2(bs))
    rank = np.array([bitrev(k % bs, bits) * ((n + bs - 1) // bs) + k // bs for k in range(n)])
    temp = np.full((n,), 1e10, np.float32)
    out = [0]
    for _ in range(1, m):
        d = sqd(xyz, xyz[out[-1]][None])
        temp = np.minimum(d, temp)
        best = temp.max()
        cand = np.nonzero(temp == best)[0]
        out.append(int(cand[np.argmin(rank[cand])]))
    return np.array(out, np.int32), temp


@pytest.mark.parametrize("n,m", [(64, 16), (128, 32), (300, 40), (1024, 64), (2048, 32)])
def test_fps_closed_form_matches_tree_simulation(oracle, n, m):
    rng = np.random.RandomState(n)
    xyz = synthetic.tie_heavy_cloud(rng, n)
    idx, temp = oracle.fps(xyz[None], m)
    bs = oracle.opt_n_threads(n)
    ref_idx, ref_temp = fps_numpy(xyz, m, bs)
    assert np.array_equal(idx[0], ref_idx)
    # temp is the state after m-1 rounds
    assert np.array_equal(temp[0], ref_temp)


def test_fps_tie_order_known_answer(oracle):
    # 8 points: point 0 at origin, points 1..7 all at the same distance from it.
    # bs = 8: thread order by bit-reversed tid: 0,4,2,6,1,5,3,7 -> among tids 1..7 the tree
    # prefers 4, then 2, 6, 1, ...; the second sample must therefore be index 4.
    xyz = np.zeros((8, 3), np.float32)
    xyz[1:, 0] = 1.0
    idx, _ = oracle.fps(xyz[None], 2)
    assert idx[0].tolist() == [0, 4]
    # with 16 points and bs=16 the preferred tid is 8
    xyz = np.zeros((16, 3), np.float32)
    xyz[1:, 0] = 1.0
    idx, _ = oracle.fps(xyz[None], 2)
    assert idx[0].tolist() == [0, 8]
    # N=24 -> bs=16, thread 8 owns k=8 only, thread 0 owns k=0,16: k=16 (tid 0) beats k=8
    xyz = np.zeros((24, 3), np.float32)
    xyz[1:, 0] = 1.0
    idx, _ = oracle.fps(xyz[None], 2)
    assert idx[0].tolist() == [0, 16]


def test_opt_n_threads_table(oracle):
    # SURVEY.md 8(a) trap 1 (values of cuda_utils.h:10-14 for the sizes on the path)
    for n, bs in [(32768, 1024), (16384, 1024), (4096, 1024), (1024, 1024), (512, 512), (256, 256), (128, 128), (300, 256), (1, 1)]:
        assert oracle.opt_n_threads(n) == bs


def ball_query_numpy(r, ns, xyz, new_xyz):
    r2 = np.float32(r) * np.float32(r)
    out = np.zeros((new_xyz.shape[0], ns), np.int32)
    for i, c in enumerate(new_xyz):
        d = sqd(c[None], xyz)
        hits = np.nonzero(d < r2)[0][:ns]
        if len(hits):
            out[i, :] = hits[0]
            out[i, :len(hits)] = hits
    return out


@pytest.mark.parametrize("r,ns", [(0.5, 16), (2.0, 32), (8.0, 8), (1e-3, 4)])
def test_ball_query_matches_numpy(oracle, r, ns):
    rng = np.random.RandomState(7)
    xyz = synthetic.tie_heavy_cloud(rng, 1500)
    new_xyz = xyz[rng.choice(1500, 200, replace=False)].copy()
    new_xyz[:5] += 100.0  # centres with no neighbour: rows stay zero
    got = oracle.ball_query(r, ns, xyz[None], new_xyz[None])[0]
    assert np.array_equal(got, ball_query_numpy(r, ns, xyz, new_xyz))
    assert (got[:5] == 0).all()


def test_three_nn_matches_numpy(oracle):
    rng = np.random.RandomState(3)
    known = synthetic.tie_heavy_cloud(rng, 257)
    unknown = synthetic.tie_heavy_cloud(rng, 999)
    d2, idx = oracle.three_nn(unknown[None], known[None])
    for i in range(0, 999, 7):
        d = sqd(unknown[i][None], known)
        order = np.lexsort((np.arange(len(d)), d))[:3]  # ascending distance, lowest index first
        assert idx[0, i].tolist() == order.tolist()
        assert np.array_equal(d2[0, i], d[order])


def test_three_nn_fewer_than_three_known(oracle):
    unknown = np.zeros((1, 4, 3), np.float32)
    known = np.ones((1, 2, 3), np.float32)
    d2, idx = oracle.three_nn(unknown, known)
    assert np.isinf(d2[0, :, 2]).all() and (idx[0, :, 2] == 0).all()
    assert (d2[0, :, :2] == 3.0).all() and idx[0, 0].tolist() == [0, 1, 0]


def test_gather_group_interpolate_shapes_and_values(oracle):
    rng = np.random.RandomState(5)
    feats = rng.randn(2, 5, 40).astype(np.float32)
    idx = rng.randint(0, 40, size=(2, 7)).astype(np.int32)
    assert np.array_equal(oracle.gather_points(feats, idx), np.take_along_axis(feats, idx[:, None, :].repeat(5, 1), 2))
    gidx = rng.randint(0, 40, size=(2, 7, 3)).astype(np.int32)
    g = oracle.group_points(feats, gidx)
    assert g.shape == (2, 5, 7, 3)
    assert np.array_equal(g[1, 2], feats[1, 2][gidx[1]])
    w = rng.rand(2, 7, 3).astype(np.float32)
    out = oracle.three_interpolate(feats, gidx, w)
    ref = (feats[0, 3][gidx[0]].astype(np.float64) * w[0]).sum(-1)
    assert np.allclose(out[0, 3], ref, rtol=1e-6, atol=1e-6)
    # backward: scatter-add is the transpose of the gather
    go = rng.randn(2, 5, 7).astype(np.float32)
    gp = oracle.gather_points_grad(go, idx, 40)
    assert np.allclose((gp * feats).sum(), (go * oracle.gather_points(feats, idx)).sum(), rtol=1e-4)
    gp3 = oracle.three_interpolate_grad(go, gidx, w, 40)
    assert np.allclose((gp3 * feats).sum(), (go * out).sum(), rtol=1e-4)


def test_nms_normal_known_answer(oracle):
    boxes = np.array([[0, 0, 2, 2, 0], [0.1, 0.1, 2.1, 2.1, 0], [5, 5, 6, 6, 0], [0, 0, 2, 2.05, 0]], np.float32)
    keep = oracle.nms_normal(boxes, 0.8)
    assert keep.tolist() == [0, 2]
    assert oracle.nms_normal(boxes, 0.99).tolist() == [0, 1, 2, 3]
    assert oracle.nms_normal(np.zeros((0, 5), np.float32), 0.5).tolist() == []


def test_roipool3d_small(oracle):
    rng = np.random.RandomState(11)
    xyz = rng.uniform(-3, 3, size=(1, 400, 3)).astype(np.float32)
    feat = rng.randn(1, 400, 4).astype(np.float32)
    boxes = np.array([[[0, 1.0, 0, 2.0, 2.0, 4.0, 0.3], [50, 0, 50, 1, 1, 1, 0.0]]], np.float32)
    pooled, empty = oracle.roipool3d(xyz, feat, boxes, sampled=16)
    assert empty.tolist() == [[0, 1]] and (pooled[0, 1] == 0).all()
    # brute force membership in float64
    c, s = np.cos(0.3), np.sin(0.3)
    dx, dz = xyz[0, :, 0] - 0, xyz[0, :, 2] - 0
    xr, zr = dx * c - dz * s, dx * s + dz * c
    inb = (np.abs(xyz[0, :, 1] - (1.0 - 1.0)) <= 1.0) & (np.abs(xr) <= 2.0) & (np.abs(zr) <= 1.0)
    ids = np.nonzero(inb)[0]
    want = [ids[k % len(ids)] for k in range(16)]
    assert np.array_equal(pooled[0, 0, :, :3], xyz[0, want])
    assert np.array_equal(pooled[0, 0, :, 3:], feat[0, want])
