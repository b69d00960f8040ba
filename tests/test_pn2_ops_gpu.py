"""GPU parity of the sm_100a pointnet2 ops, called through the reference-shaped API
(pointnet2_utils -> pointnet2_cuda -> C-ABI), against
  (1) the CPU oracle on seeded inputs at sizes it finishes in seconds, and
  (2) the reference's own kernels (oracle/_ref/libpn2_legacy.so) at BASELINE.json's full sizes.
Index outputs and the f32 values that are pure copies / fixed-order FMAs must be bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load

pytestmark = pytest.mark.gpu
synthetic = load("synthetic")


def p2u():
    return load("pointnet2_utils")


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
@pytest.mark.parametrize("n,m,b", [(2048, 512, 2), (512, 128, 5), (128, 32, 3), (300, 77, 2), (4096, 256, 1)])
def test_fps_vs_oracle(cuda, oracle, kind, n, m, b):
    xyz_h = synthetic.make_clouds(kind, b, n, seed=1024 + n)
    idx = p2u().furthest_point_sample(torch.from_numpy(xyz_h).to(cuda), m)
    ref, _ = oracle.fps(xyz_h, m)
    assert idx.dtype == torch.int32 and idx.shape == (b, m)
    assert np.array_equal(idx.cpu().numpy(), ref)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_fps_every_cluster_size_and_temp_writeback(cuda, oracle, cluster):
    cabi = load("cabi")
    p2c = load("pointnet2_cuda")
    xyz_h = synthetic.make_clouds("ties", 3, 3000, seed=666)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    temp = torch.full((3, 3000), 1e10, device=cuda)
    idx = torch.empty((3, 200), dtype=torch.int32, device=cuda)
    cabi.call("pn2_fps_cluster_f32", cabi.ptr(xyz), cabi.ptr(temp), cabi.ptr(idx), cabi.i32(3), cabi.i32(3000), cabi.i32(200),
              cabi.i32(cluster))
    ref, ref_temp = oracle.fps(xyz_h, 200)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(temp.cpu().numpy(), ref_temp)  # caller scratch mutated like the reference


@pytest.mark.parametrize("warps", [0, 4, 8, 16])
@pytest.mark.parametrize("kind,n,m,b", [("lidar", 16384, 1024, 2), ("ties", 16384, 700, 2), ("uniform", 9000, 600, 2),
                                         ("lidar", 6000, 500, 2),
                                         ("ties", 4096, 512, 3), ("lidar", 3000, 3000, 1), ("ties", 2049, 300, 2),
                                         ("lidar", 130, 64, 3), ("ties", 64, 64, 2)])
def test_fps_cells_kernel_vs_oracle(cuda, oracle, warps, kind, n, m, b):
    """csrc/fps_cells.cu (exact spatial pruning, one CTA per cloud) for every CTA size: indices AND the running
    min-distances it leaves in the caller's scratch equal the C restatement of sampling_gpu.cu:93-209."""
    cabi = load("cabi")
    xyz_h = synthetic.make_clouds(kind, b, n, seed=77 + n)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    temp = torch.full((b, n), 1e10, device=cuda)
    idx = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
    cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(temp), cabi.ptr(idx), cabi.i32(b), cabi.i32(n), cabi.i32(m),
              cabi.i32(warps))
    ref, ref_temp = oracle.fps(xyz_h, m)
    assert np.array_equal(idx.cpu().numpy(), ref), (kind, n, m, int((idx.cpu().numpy() != ref).sum()))
    assert np.array_equal(temp.cpu().numpy(), ref_temp)
    # without the scratch (the inference path) and through the heuristic of pn2_fps_f32
    idx2 = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
    cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx2), cabi.i32(b), cabi.i32(n), cabi.i32(m),
              cabi.i32(warps))
    assert torch.equal(idx2, idx)


@pytest.mark.parametrize("perm", [(1, 0, 2), (2, 1, 0), (0, 2, 1)])
def test_fps_cells_grid_follows_the_widest_axes(cuda, oracle, perm):
    """The Hilbert grid of csrc/fps_cells.cu spans the two widest axes of the cloud, whichever they are (a sweep whose
    flat axis is not y, a wall, a line): the pruning changes, the indices never do."""
    cabi = load("cabi")
    base = synthetic.make_clouds("lidar", 2, 6000, seed=9)[:, :, list(perm)].copy()
    line = np.zeros((1, 6000, 3), np.float32)
    line[0, :, perm[0]] = np.random.RandomState(3).uniform(-5, 5, 6000).astype(np.float32)
    for xyz_h in (base, np.ascontiguousarray(line)):
        b, n, m = xyz_h.shape[0], xyz_h.shape[1], 400
        xyz = torch.from_numpy(xyz_h).to(cuda)
        idx = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
        cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(b), cabi.i32(n), cabi.i32(m),
                  cabi.i32(0))
        ref, _ = oracle.fps(xyz_h, m)
        assert np.array_equal(idx.cpu().numpy(), ref)


def test_fps_cells_rejects_what_it_cannot_hold(cuda):
    cabi = load("cabi")
    xyz = torch.zeros((1, 20000, 3), device=cuda)
    idx = torch.empty((1, 8), dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError):
        cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(1), cabi.i32(20000), cabi.i32(8),
                  cabi.i32(0))
    with pytest.raises(RuntimeError):
        cabi.call("pn2_fps_cells_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(1), cabi.i32(4096), cabi.i32(8),
                  cabi.i32(5))


@pytest.mark.parametrize("kind,n,m", [("lidar", 512, 128), ("ties", 512, 128), ("uniform", 128, 32), ("ties", 1000, 100),
                                      ("lidar", 65, 65)])
def test_fps_many_small_clouds_take_the_two_warp_kernel(cuda, oracle, kind, n, m):
    """b >= 256 clouds of at most 1024 points (the RCNN stage: 1600 ROI clouds) run fps_kernel<P, 1, 64>: indices and the
    scratch left behind equal the oracle's, like every other launch shape."""
    b = 300
    xyz_h = synthetic.make_clouds(kind, b, n, seed=n + m)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    temp = torch.full((b, n), 1e10, device=cuda)
    idx = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
    load("pointnet2_cuda").furthest_point_sampling_wrapper(b, n, m, xyz, temp, idx)
    ref, ref_temp = oracle.fps(xyz_h, m)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(temp.cpu().numpy(), ref_temp)


def test_fps_edge_cases(cuda, oracle):
    f = p2u().furthest_point_sample
    one = torch.zeros((2, 1, 3), device=cuda)
    assert f(one, 1).cpu().tolist() == [[0], [0]]
    same = torch.ones((1, 64, 3), device=cuda)  # all points identical: every distance ties at 0
    ref, _ = oracle.fps(same.cpu().numpy(), 8)
    assert np.array_equal(f(same, 8).cpu().numpy(), ref)
    assert f(torch.zeros((0, 16, 3), device=cuda), 4).shape == (0, 4)


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
def test_fps_full_size_vs_legacy(cuda, legacy, kind):
    # config 3 of BASELINE.json: 16384 -> 4096, batch 8 ; plus the RCNN shapes 512->128, 128->32
    for b, n, m in [(8, 16384, 4096), (64, 512, 128), (64, 128, 32), (2, 32768, 1024)]:
        xyz = torch.from_numpy(synthetic.make_clouds(kind, b, n, seed=1024)).to(cuda)
        got = p2u().furthest_point_sample(xyz, m)
        ref, _ = legacy.fps(xyz, m)
        assert torch.equal(got, ref), (kind, b, n, m, int((got != ref).sum()))


@pytest.mark.parametrize("r,ns", [(0.1, 16), (0.5, 32), (1.0, 64), (4.0, 16)])
def test_ball_query_vs_oracle(cuda, oracle, r, ns):
    xyz_h = synthetic.make_clouds("lidar", 2, 3000, seed=5)
    new_h = xyz_h[:, ::7].copy()
    new_h[:, :3] += 500.0  # no neighbour at all -> rows stay zero
    got = p2u().ball_query(r, ns, torch.from_numpy(xyz_h).to(cuda), torch.from_numpy(new_h).to(cuda))
    assert np.array_equal(got.cpu().numpy(), oracle.ball_query(r, ns, xyz_h, new_h))


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
def test_ball_query_full_size_vs_legacy_and_dual(cuda, legacy, kind):
    cabi = load("cabi")
    xyz = torch.from_numpy(synthetic.make_clouds(kind, 4, 16384, seed=1024)).to(cuda)
    idx, _ = legacy.fps(xyz, 4096)
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    refs = {}
    for r, ns in [(0.1, 16), (0.5, 32), (0.1, 64)]:
        refs[(r, ns)] = legacy.ball_query(r, ns, xyz, new_xyz)
        assert torch.equal(p2u().ball_query(r, ns, xyz, new_xyz), refs[(r, ns)]), (kind, r, ns)
    i0 = torch.zeros((4, 4096, 16), dtype=torch.int32, device=cuda)
    i1 = torch.zeros((4, 4096, 32), dtype=torch.int32, device=cuda)
    cabi.call("pn2_ball_query_dual_f32", cabi.ptr(new_xyz), cabi.ptr(xyz), cabi.ptr(i0), cabi.ptr(i1), cabi.i32(4),
              cabi.i32(16384), cabi.i32(4096), cabi.f32(0.1), cabi.i32(16), cabi.f32(0.5), cabi.i32(32))
    assert torch.equal(i0, refs[(0.1, 16)]) and torch.equal(i1, refs[(0.5, 32)])


def _culled(cabi, xyz, new_xyz, r0, ns0, r1=0.0, ns1=0, order=True):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    i0 = torch.zeros((B, M, ns0), dtype=torch.int32, device=xyz.device)
    i1 = torch.zeros((B, M, max(ns1, 1)), dtype=torch.int32, device=xyz.device)
    scratch = torch.empty((B, M), dtype=torch.int32, device=xyz.device) if order else None
    cabi.call("pn2_ball_query_culled_f32", cabi.ptr(new_xyz), cabi.ptr(xyz), cabi.ptr(i0), cabi.ptr(i1 if ns1 else None),
              cabi.ptr(scratch), cabi.i32(B), cabi.i32(N), cabi.i32(M), cabi.f32(r0), cabi.i32(ns0), cabi.f32(r1),
              cabi.i32(ns1))
    return i0, i1, scratch


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
@pytest.mark.parametrize("n,m", [(3000, 700), (1024, 256), (4096, 1024), (300, 77), (256, 64), (129, 1)])
def test_ball_query_culled_vs_oracle(cuda, oracle, kind, n, m):
    """the Hilbert-ordered, ballot-compacted scan returns the reference's idx bit for bit"""
    cabi = load("cabi")
    xyz_h = synthetic.make_clouds(kind, 2, n, seed=77 + n)
    sel = np.random.RandomState(n).permutation(n)[:m]
    new_h = xyz_h[:, sel].copy()
    new_h[:, :min(5, m - 1)] += 500.0     # centres without any neighbour: rows stay zero
    xyz, new_xyz = torch.from_numpy(xyz_h).to(cuda), torch.from_numpy(new_h).to(cuda)
    for (r0, ns0, r1, ns1) in [(0.1, 16, 0.5, 32), (1.0, 16, 2.0, 32), (0.5, 64, 0.0, 0), (100.0, 8, 0.05, 4)]:
        i0, i1, order = _culled(cabi, xyz, new_xyz, r0, ns0, r1, ns1)
        assert np.array_equal(i0.cpu().numpy(), oracle.ball_query(r0, ns0, xyz_h, new_h)), (kind, n, m, r0)
        if ns1:
            assert np.array_equal(i1.cpu().numpy(), oracle.ball_query(r1, ns1, xyz_h, new_h)), (kind, n, m, r1)
        if m >= 256:   # the scratch holds a permutation of the centres of every cloud
            assert torch.equal(torch.sort(order, dim=1)[0], torch.arange(m, device=cuda, dtype=torch.int32).expand(2, m))


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
def test_ball_query_culled_full_size_vs_brute_force(cuda, legacy, kind):
    cabi = load("cabi")
    xyz = torch.from_numpy(synthetic.make_clouds(kind, 4, 16384, seed=1024)).to(cuda)
    idx, _ = legacy.fps(xyz, 4096)
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    i0, i1, _ = _culled(cabi, xyz, new_xyz, 0.1, 16, 0.5, 32)
    assert torch.equal(i0, legacy.ball_query(0.1, 16, xyz, new_xyz))
    assert torch.equal(i1, legacy.ball_query(0.5, 32, xyz, new_xyz))
    # no scratch -> brute-force path, same answer; fused.ball_query_dual is the product caller
    j0, j1, _ = _culled(cabi, xyz, new_xyz, 0.1, 16, 0.5, 32, order=False)
    assert torch.equal(i0, j0) and torch.equal(i1, j1)
    k0, k1 = load("fused").ball_query_dual(xyz, new_xyz, 0.1, 16, 0.5, 32)
    assert torch.equal(i0, k0) and torch.equal(i1, k1)
    # the reference-shaped single-radius C entry (brute force) and the wrapper (culled) agree as well
    s0 = torch.zeros_like(i1)
    cabi.call("pn2_ball_query_f32", cabi.ptr(new_xyz), cabi.ptr(xyz), cabi.ptr(s0), cabi.i32(4), cabi.i32(16384),
              cabi.i32(4096), cabi.f32(0.5), cabi.i32(32))
    assert torch.equal(s0, i1) and torch.equal(p2u().ball_query(0.5, 32, xyz, new_xyz), i1)
    # RCNN shape: many small clouds, one radius, 64 samples
    small = torch.from_numpy(synthetic.make_clouds(kind, 50, 512, seed=3)).to(cuda) * 0.05
    cen = small[:, ::4].contiguous()
    ref = torch.zeros((50, 128, 64), dtype=torch.int32, device=cuda)
    cabi.call("pn2_ball_query_f32", cabi.ptr(cen), cabi.ptr(small), cabi.ptr(ref), cabi.i32(50), cabi.i32(512),
              cabi.i32(128), cabi.f32(0.2), cabi.i32(64))
    assert torch.equal(p2u().ball_query(0.2, 64, small, cen), ref)
    # degenerate cloud: every point identical (zero-size bounding box)
    same = torch.ones((1, 2048, 3), device=cuda)
    d0, _, _ = _culled(cabi, same, same[:, :512].contiguous(), 0.1, 16)
    assert torch.equal(d0, torch.arange(16, device=cuda, dtype=torch.int32).expand(1, 512, 16))


def test_gather_group_vs_oracle_and_backward(cuda, oracle):
    rng = np.random.RandomState(0)
    feats_h = rng.randn(3, 19, 700).astype(np.float32)
    idx_h = rng.randint(0, 700, size=(3, 130)).astype(np.int32)
    gidx_h = rng.randint(0, 700, size=(3, 130, 9)).astype(np.int32)
    feats = torch.from_numpy(feats_h).to(cuda).requires_grad_(True)
    out = p2u().gather_operation(feats, torch.from_numpy(idx_h).to(cuda))
    assert np.array_equal(out.detach().cpu().numpy(), oracle.gather_points(feats_h, idx_h))
    g = p2u().grouping_operation(feats, torch.from_numpy(gidx_h).to(cuda))
    assert np.array_equal(g.detach().cpu().numpy(), oracle.group_points(feats_h, gidx_h))
    go_h = rng.randn(*g.shape).astype(np.float32)
    g.backward(torch.from_numpy(go_h).to(cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), oracle.group_points_grad(go_h, gidx_h, 700), rtol=1e-5, atol=1e-5)
    feats.grad = None
    go2 = rng.randn(*out.shape).astype(np.float32)
    out.backward(torch.from_numpy(go2).to(cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), oracle.gather_points_grad(go2, idx_h, 700), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,m", [(1024, 256), (999, 5), (64, 2), (4096, 1024)])
def test_three_nn_and_interpolate_vs_oracle(cuda, oracle, n, m):
    known_h = synthetic.make_clouds("ties", 2, m, seed=n)
    unknown_h = synthetic.make_clouds("ties", 2, n, seed=n + 1)
    dist, idx = p2u().three_nn(torch.from_numpy(unknown_h).to(cuda), torch.from_numpy(known_h).to(cuda))
    d2_ref, idx_ref = oracle.three_nn(unknown_h, known_h)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(d2_ref))  # the Python wrapper returns sqrt (pointnet2_utils.py:98)
    rng = np.random.RandomState(1)
    feats_h = rng.randn(2, 21, m).astype(np.float32)
    w_h = rng.rand(2, n, 3).astype(np.float32)
    feats = torch.from_numpy(feats_h).to(cuda).requires_grad_(True)
    out = p2u().three_interpolate(feats, idx, torch.from_numpy(w_h).to(cuda))
    assert np.array_equal(out.detach().cpu().numpy(), oracle.three_interpolate(feats_h, idx_ref, w_h))
    go = rng.randn(*out.shape).astype(np.float32)
    out.backward(torch.from_numpy(go).to(cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), oracle.three_interpolate_grad(go, idx_ref, w_h, m), rtol=1e-4, atol=1e-4)


def test_three_nn_interpolate_full_size_vs_legacy(cuda, legacy):
    # FP0 of the RPN backbone: 16384 unknown <- 4096 known, 256 channels
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=1024)).to(cuda)
    known = xyz[:, :4096].contiguous()
    dist, idx = p2u().three_nn(xyz, known)
    d2_ref, idx_ref = legacy.three_nn(xyz, known)
    assert torch.equal(idx, idx_ref) and torch.equal(dist, torch.sqrt(d2_ref))
    feats = torch.randn((2, 256, 4096), device=cuda)
    w = torch.rand((2, 16384, 3), device=cuda)
    assert torch.equal(p2u().three_interpolate(feats, idx, w), legacy.three_interpolate(feats, idx, w))
    assert torch.equal(p2u().grouping_operation(feats, idx), legacy.group(feats, idx))
    assert torch.equal(p2u().gather_operation(feats, idx[:, :, 0].contiguous()), legacy.gather(feats, idx[:, :, 0].contiguous()))


@pytest.mark.parametrize("kind", ["uniform", "lidar", "ties"])
@pytest.mark.parametrize("n,m,b", [(16384, 4096, 3), (4096, 1024, 2), (2001, 601, 2)])
def test_three_nn_culled_equals_brute_force(cuda, kind, n, m, b):
    """Hilbert-ordered, bound-culled three_nn: dist2 and idx (ties included) identical to the full scan"""
    cabi = load("cabi")
    unknown = torch.from_numpy(synthetic.make_clouds(kind, b, n, seed=9 + n)).to(cuda)
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(m))[:m].to(cuda)
    known = unknown[:, sel].contiguous()          # known points coincide with unknown ones: zero distances and ties
    outs = []
    for culled in (False, True):
        d2 = torch.empty((b, n, 3), device=cuda)
        idx = torch.empty((b, n, 3), dtype=torch.int32, device=cuda)
        order = torch.empty((b, n), dtype=torch.int32, device=cuda) if culled else None
        cabi.call("pn2_three_nn_culled_f32", cabi.ptr(unknown), cabi.ptr(known), cabi.ptr(d2), cabi.ptr(idx),
                  cabi.ptr(order), cabi.i32(b), cabi.i32(n), cabi.i32(m))
        outs.append((d2, idx))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(torch.sort(order, dim=1)[0], torch.arange(n, device=cuda, dtype=torch.int32).expand(b, n))


def test_legacy_pins_the_oracle(cuda, legacy, oracle):
    """The CPU restatement agrees with the reference's real kernels (the oracle's pin)."""
    for kind in ["uniform", "lidar", "ties"]:
        xyz_h = synthetic.make_clouds(kind, 2, 4096, seed=1024)
        xyz = torch.from_numpy(xyz_h).to(cuda)
        idx, temp = legacy.fps(xyz, 512)
        ref_idx, ref_temp = oracle.fps(xyz_h, 512)
        assert np.array_equal(idx.cpu().numpy(), ref_idx) and np.array_equal(temp.cpu().numpy(), ref_temp)
        new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        for r, ns in [(0.5, 16), (1.0, 32)]:
            assert np.array_equal(legacy.ball_query(r, ns, xyz, new_xyz).cpu().numpy(),
                                  oracle.ball_query(r, ns, xyz_h, new_xyz.cpu().numpy()))
        d2, i3 = legacy.three_nn(xyz, new_xyz)
        d2o, i3o = oracle.three_nn(xyz_h, new_xyz.cpu().numpy())
        assert np.array_equal(i3.cpu().numpy(), i3o) and np.array_equal(d2.cpu().numpy(), d2o)


@pytest.mark.parametrize("kind,n0,n1,m", [("lidar", 16384, 4096, 1024), ("lidar", 4096, 1024, 256), ("uniform", 1024, 256, 64),
                                           ("ties", 8192, 2048, 512), ("ties", 600, 300, 150), ("lidar", 512, 128, 32)])
def test_fps_prefix_check_and_guarded_launch(cuda, oracle, kind, n0, n1, m):
    """pn2_fps_prefix_check_f32 + pn2_fps_guarded_f32 (the backbone's levels 2-4 and RCNN SA2) against the oracle:
    the sampled cloud is the FPS-ordered output of a previous level.  Where the test passes the kernel writes arange(m)
    and the oracle's sequential answer must be exactly that; `ties` clouds (lattice points, exact distance ties) make the
    test fail for some clouds, which then take the real round loop -- bit-exact either way."""
    fz = load("fused")
    B = 4
    xyz0_h = synthetic.make_clouds(kind, B, n0, seed=n0 + m)
    xyz0 = torch.from_numpy(xyz0_h).to(cuda)
    _, xyz1 = fz.fps_gather(xyz0, n1)                               # level l-1 (the real kernel)
    ref1, _ = oracle.fps(xyz0_h, n1)
    xyz1_h = xyz1.cpu().numpy()
    assert np.array_equal(xyz1_h, np.take_along_axis(xyz0_h, ref1[:, :, None].astype(np.int64), axis=1))
    viol = torch.zeros((B,), dtype=torch.int32, device=cuda)
    dmin = torch.empty((B, m), dtype=torch.float32, device=cuda)
    cabi = load("cabi")
    cabi.call("pn2_fps_prefix_check_f32", cabi.ptr(xyz1), cabi.ptr(dmin), cabi.ptr(viol), cabi.i32(B), cabi.i32(n1), cabi.i32(m))
    idx, _ = fz.fps_gather(xyz1, m, fps_ordered=True)               # check + guarded launch
    ref2, _ = oracle.fps(xyz1_h, m)
    assert np.array_equal(idx.cpu().numpy(), ref2)
    v = viol.cpu().numpy()
    for b in range(B):
        if v[b] == 0:
            assert np.array_equal(ref2[b], np.arange(m)), "prefix test passed but the reference answer is not arange"
    print("%s %d->%d->%d: %d of %d clouds skip the round loop" % (kind, n0, n1, m, int((v == 0).sum()), B))
    if kind == "lidar":
        assert (v == 0).all()


def test_fps_prefix_check_rejects_unordered_and_duplicate_clouds(cuda, oracle):
    """an arbitrary (not FPS-ordered) cloud and a cloud with duplicated points must fail the test and still sample exactly"""
    fz = load("fused")
    xyz_h = synthetic.make_clouds("uniform", 3, 2048, seed=11)
    xyz_h[1, 7] = xyz_h[1, 3]                                       # duplicate of an early point
    xyz = torch.from_numpy(xyz_h).to(cuda)
    viol = torch.zeros((3,), dtype=torch.int32, device=cuda)
    dmin = torch.empty((3, 256), dtype=torch.float32, device=cuda)
    cabi = load("cabi")
    cabi.call("pn2_fps_prefix_check_f32", cabi.ptr(xyz), cabi.ptr(dmin), cabi.ptr(viol), cabi.i32(3), cabi.i32(2048), cabi.i32(256))
    assert (viol.cpu().numpy() > 0).all()
    idx, _ = fz.fps_gather(xyz, 256, fps_ordered=True)
    ref, _ = oracle.fps(xyz_h, 256)
    assert np.array_equal(idx.cpu().numpy(), ref)


@pytest.mark.parametrize("n,m,b", [(16384, 512, 2), (3000, 3000, 1), (512, 128, 5), (20000, 64, 2), (1, 1, 3)])
def test_fps_writes_the_coordinates_of_its_picks(cuda, n, m, b):
    """pn2_fps_xyz_f32 / pn2_fps_guarded_xyz_f32 (pruned one-CTA kernel, plain kernel, cluster kernel, guarded exit):
    new_xyz is exactly xyz gathered at the indices, and the indices are those of pn2_fps_f32."""
    cabi = load("cabi")
    xyz_h = synthetic.make_clouds("ties" if n == 3000 else "lidar", b, n, seed=5 + n)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    ref = p2u().furthest_point_sample(xyz, m)
    idx = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
    new_xyz = torch.full((b, m, 3), float("nan"), device=cuda)
    cabi.call("pn2_fps_xyz_f32", cabi.ptr(xyz), cabi.ptr(idx), cabi.ptr(new_xyz), cabi.i32(b), cabi.i32(n), cabi.i32(m))
    assert torch.equal(idx, ref)
    assert torch.equal(new_xyz, torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)))
    # guarded: cloud 0 flagged "provably arange" (only true for an FPS-ordered cloud: reorder it first), the others not
    ordered = torch.gather(xyz, 1, ref.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()       # (b, m, 3), FPS order
    mm = max(m // 2, 1)
    viol = torch.empty((b,), dtype=torch.int32, device=cuda)
    dmin = torch.empty((b, mm), device=cuda)
    if mm <= 4096:
        cabi.call("pn2_fps_prefix_check_f32", cabi.ptr(ordered), cabi.ptr(dmin), cabi.ptr(viol), cabi.i32(b), cabi.i32(m), cabi.i32(mm))
        idx2 = torch.full((b, mm), -1, dtype=torch.int32, device=cuda)
        new2 = torch.full((b, mm, 3), float("nan"), device=cuda)
        cabi.call("pn2_fps_guarded_xyz_f32", cabi.ptr(ordered), cabi.ptr(idx2), cabi.ptr(new2), cabi.ptr(viol), cabi.i32(b),
                  cabi.i32(m), cabi.i32(mm))
        assert torch.equal(idx2, p2u().furthest_point_sample(ordered, mm))
        assert torch.equal(new2, torch.gather(ordered, 1, idx2.long().unsqueeze(-1).expand(-1, -1, 3)))


@pytest.mark.parametrize("n,m", [(16384, 4096), (4096, 1024), (300, 40), (100, 7)])
def test_ball_query_fill_variant_needs_no_zeroed_lists(cuda, n, m):
    """pn2_ball_query_culled_fill_f32 on garbage-initialised lists == zero fill + pn2_ball_query_culled_f32, including
    centres without a single neighbour (single and dual radius, culled and brute-force sizes)."""
    cabi = load("cabi")
    xyz_h = synthetic.make_clouds("lidar", 2, n, seed=31)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    new_xyz = xyz[:, :m].clone()
    new_xyz[:, ::5] += 300.0                                  # every fifth centre has no neighbour at all
    for r0, ns0, r1, ns1 in ((0.1, 16, 0.5, 32), (0.4, 64, 0.0, 0)):
        outs = []
        for name, init in (("pn2_ball_query_culled_f32", 0), ("pn2_ball_query_culled_fill_f32", -7)):
            i0 = torch.full((2, m, ns0), init, dtype=torch.int32, device=cuda)
            i1 = torch.full((2, m, max(ns1, 1)), init, dtype=torch.int32, device=cuda)
            order = torch.empty((2, m), dtype=torch.int32, device=cuda)
            hits = torch.full((2, 2, m), -5, dtype=torch.int32, device=cuda)
            extra = (cabi.ptr(hits[0]), cabi.ptr(hits[1] if ns1 else None)) if "fill" in name else ()
            cabi.call(name, cabi.ptr(new_xyz), cabi.ptr(xyz), cabi.ptr(i0), cabi.ptr(i1 if ns1 else None), cabi.ptr(order),
                      cabi.i32(2), cabi.i32(n), cabi.i32(m), cabi.f32(r0), cabi.i32(ns0), cabi.f32(r1), cabi.i32(ns1), *extra)
            outs.append((i0, i1 if ns1 else None, hits))
        assert torch.equal(outs[0][0], outs[1][0])
        assert int((outs[1][0][:, ::5] != 0).sum()) == 0
        if ns1:
            assert torch.equal(outs[0][1], outs[1][1])
        # hit counts: the neighbours found, capped at nsample = the number of distinct entries of a list (0 for an empty one)
        for lst, h in ((outs[1][0], outs[1][2][0]),) + (((outs[1][1], outs[1][2][1]),) if ns1 else ()):
            distinct = (lst[:, :, 1:] != lst[:, :, :1]).sum(dim=2) + 1
            found = torch.where(new_xyz[:, :, 0] > 200.0, torch.zeros_like(distinct), distinct)      # the shifted centres: none
            assert torch.equal(h.long(), found)


def test_three_interpolate_from_squared_distances_is_the_torch_weighting(cuda):
    """pn2_three_interpolate_pm_d2_f32 == sqrt / reciprocal / sum / divide in torch (pointnet2_utils.py:104,
    pointnet2_modules.py:209-211) + pn2_three_interpolate_pm_f32, bit for bit -- including coincident points (d2 = 0)."""
    fz = load("fused")
    b, n, m, c = 2, 5000, 1200, 96
    unknown = torch.from_numpy(synthetic.make_clouds("lidar", b, n, seed=2)).to(cuda)
    known = unknown[:, :m].contiguous()                      # the first m unknown points coincide with a known one
    feats = torch.randn((b, m, c), device=cuda)
    dist, idx = p2u().three_nn(unknown, known)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = (dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)).contiguous()
    ref = torch.empty((b * n, c), device=cuda)
    fz.three_interpolate_pm(feats, idx, weight, ref)
    dist2 = torch.empty((b, n, 3), device=cuda)
    idx2 = torch.empty((b, n, 3), dtype=torch.int32, device=cuda)
    load("pointnet2_cuda").three_nn_wrapper(b, n, m, unknown, known, dist2, idx2)
    assert torch.equal(idx, idx2)
    same = []
    for order in (0, 1, 2):
        out = torch.empty((b * n, c), device=cuda)
        fz.three_interpolate_pm_d2(feats, idx2, dist2, out, sum_order=order)
        same.append(bool(torch.equal(out, ref)))
    assert same[fz.INTERP_SUM_ORDER], ("association that matches torch.sum: %s, configured: %d" % (same, fz.INTERP_SUM_ORDER))


def test_group_compaction_from_the_ball_querys_hit_counts(cuda):
    """fused.ball_query_single / _dual leave the per-centre hit counts with the lists; group_compact() then skips its counting pass
    over the lists (pn2_group_compact_lists_i32 with hits): same compact lists and row count as counting from the lists."""
    fz = load("fused")
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 3, 6000, seed=12)).to(cuda)
    centres = xyz[:, :700].clone()
    centres[:, ::7] += 300.0                                  # centres without any neighbour
    lists = list(fz.ball_query_dual(xyz, centres, 0.3, 16, 1.2, 32)) + [fz.ball_query_single(xyz, centres, 0.6, 64)]
    for idx in lists:
        assert getattr(idx, "_pn2_hits", None) is not None
        got = fz.group_compact(idx, align=8)
        saved = fz.COMPACT_USE_HITS
        try:
            fz.COMPACT_USE_HITS = False
            want = fz.group_compact(idx, align=8)
        finally:
            fz.COMPACT_USE_HITS = saved
        u = int(want[2].item())
        assert int(got[2].item()) == u
        assert torch.equal(got[0][:u], want[0][:u]) and torch.equal(got[1][:u], want[1][:u])
