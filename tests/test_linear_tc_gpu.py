"""GPU parity of the tcgen05 shared-MLP kernels (csrc/linear_tc.cu: BF16x3 split operands, fp32
accumulation in TMEM) against an fp64 PyTorch reference of the same op.  Tolerance: 1e-4
relative to the tensor scale (BASELINE.json north_star: "within 1e-4 rel for MLP/feature
tensors"); the measured error of one layer is ~7e-6 of the scale."""
import numpy as np
import pytest
import torch

from conftest import load

pytestmark = pytest.mark.gpu
synthetic = load("synthetic")


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _layer(fz, cout, cin, relu, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.randn((cout, cin), generator=g) / cin ** 0.5).cuda()
    b = torch.randn((cout,), generator=g).cuda()
    return w, b, fz.PackedLayerTC(w, b, relu)


@pytest.mark.parametrize("rows,cin,cout,relu", [
    (128, 64, 128, True), (128, 16, 16, False), (1000, 3, 16, True), (4096, 131, 128, True), (777, 256, 46, False),
    (300, 1536, 512, True), (5, 7, 1, False), (2048, 96, 130, True), (20000, 128, 256, True), (50000, 128, 76, False),
    (999, 515, 384, True), (640, 259, 196, True)])
def test_linear_tc_vs_fp64(cuda, rows, cin, cout, relu):
    fz = load("fused")
    w, b, layer = _layer(fz, cout, cin, relu, rows + cin)
    x = torch.randn((rows, cin), generator=torch.Generator(device="cpu").manual_seed(1)).cuda()
    y = fz.linear_tc(x, layer)
    ref = x.double() @ w.double().t() + b.double()
    ref = ref.clamp_min(0) if relu else ref
    e = rel_err(y, ref)
    assert e < 1e-4, e
    assert e < 3e-5, "BF16x3 should be ~1e-5: %g" % e


def test_linear_tc_strided_residual_pool(cuda):
    fz = load("fused")
    w, b, layer = _layer(fz, 64, 100, True, 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    big = torch.randn((4096, 200), generator=g).cuda()
    res = torch.randn((4096, 64), generator=g).cuda()
    x = big[:, 50:150]                                     # unaligned column slice, ld 200
    out = torch.zeros((4096, 96), device=cuda)
    fz.linear_tc(x, layer, out=out[:, 32:], res=res)
    ref = torch.relu(x.double() @ w.double().t() + b.double() + res.double())
    assert rel_err(out[:, 32:], ref) < 3e-5
    assert float(out[:, :32].abs().max()) == 0.0
    for pool in (16, 32, 64, 128):
        y = fz.linear_tc(x, layer, pool=pool)
        r = torch.relu(x.double() @ w.double().t() + b.double()).view(4096 // pool, pool, 64).max(1)[0]
        assert y.shape == r.shape and rel_err(y, r) < 3e-5, pool
    # ragged tail: rows not a multiple of 128
    y = fz.linear_tc(x[:1000 * 1], layer, pool=1)
    assert rel_err(y, torch.relu(x[:1000].double() @ w.double().t() + b.double())) < 3e-5
    y = fz.linear_tc(x[:64 * 9], layer, pool=64)
    r = torch.relu(x[:576].double() @ w.double().t() + b.double()).view(9, 64, 64).max(1)[0]
    assert rel_err(y, r) < 3e-5


def test_linear_cat_two_sources_equals_concatenated(cuda):
    """pn2_linear_tc2_f32 reads [xa | xb] from two strided sources; same K order as the one-source
    kernel on the materialised concatenation, so the results are identical bit for bit."""
    fz = load("fused")
    fz.set_mlp_engine("tc")
    w, b, _ = _layer(fz, 128, 256, True, 3)
    layer = fz.PackedLayer(w, b, True)
    g = torch.Generator(device="cpu").manual_seed(1)
    for rows in (128 * 40, 1000):
        xa = torch.randn((rows, 128), generator=g).cuda()
        wide = torch.randn((rows, 136), generator=g).cuda()       # padded pooled rows: features at column 8
        xb = wide[:, 8:]
        got = fz.linear_cat(xa, xb, layer)
        cat = torch.cat((xa, xb), dim=1)
        assert torch.equal(got, fz.linear(cat, layer))
        assert rel_err(got, torch.relu(cat.double() @ w.double().t() + b.double())) < 3e-5
    # residual epilogue (prefetched loads) on full and ragged tiles
    w2, b2, l2 = _layer(fz, 64, 128, True, 4)
    for rows in (128 * 33, 777):
        x = torch.randn((rows, 128), generator=g).cuda()
        res = torch.randn((rows, 64), generator=g).cuda()
        got = fz.linear_tc(x, l2, res=res)
        assert rel_err(got, torch.relu(x.double() @ w2.double().t() + b2.double() + res.double())) < 3e-5


@pytest.mark.parametrize("c1,cout,ns,pool", [(128, 128, 64, 1), (128, 256, 64, 64), (16, 16, 16, 1), (32, 64, 32, 32), (196, 256, 16, 16)])
def test_sa_group_linear_tc_vs_fp64(cuda, c1, cout, ns, pool):
    fz = load("fused")
    B, N, M = 3, 512, 96
    g = torch.Generator(device="cpu").manual_seed(c1 + cout)
    xyz = torch.rand((B, N, 3), generator=g).cuda()
    centres = xyz[:, :M].contiguous()
    idx = torch.randint(0, N, (B, M, ns), generator=g, dtype=torch.int32).cuda()
    h = torch.randn((B * N, c1), generator=g).cuda()
    wxyz = torch.randn((3, c1), generator=g).cuda()
    w, b, layer = _layer(fz, cout, c1, True, 7)
    y = fz.sa_group_linear_tc(h, idx, xyz, centres, wxyz, layer, pool=pool)
    # fp64 reference of: a = relu(H[j] + d . Wx) ; y = relu(a W^T + b) ; max over ns
    j = idx.long()
    hj = h.view(B, N, c1).double()[torch.arange(B).view(B, 1, 1), j]                       # (B,M,ns,c1)
    d = xyz.double()[torch.arange(B).view(B, 1, 1), j] - centres.double().unsqueeze(2)      # (B,M,ns,3)
    a = torch.relu(hj + d @ wxyz.double())
    ref = torch.relu(a @ w.double().t() + b.double()).view(B * M * ns, cout)
    if pool > 1:
        ref = ref.view(B * M * ns // pool, pool, cout).max(1)[0]
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 3e-5


def _sa_fused_case(cuda, fz, c1, c2, c3, ns, B, N, M):
    g = torch.Generator(device="cpu").manual_seed(c1 + c3 + B)
    xyz = torch.rand((B, N, 3), generator=g).cuda()
    centres = xyz[:, :M].contiguous()
    idx = torch.randint(0, N, (B, M, ns), generator=g, dtype=torch.int32).cuda()
    h = torch.randn((B * N, c1), generator=g).cuda()
    wxyz = torch.randn((3, c1), generator=g).cuda()
    w2 = (torch.randn((c2, c1), generator=g) / c1 ** 0.5).cuda(); b2 = torch.randn((c2,), generator=g).cuda()
    w3 = (torch.randn((c3, c2), generator=g) / c2 ** 0.5).cuda(); b3 = torch.randn((c3,), generator=g).cuda()
    l2, l3 = fz.PackedLayer(w2, b2, True), fz.PackedLayer(w3, b3, True)
    out = torch.full((B * M, c3 + 8), -1.0, device=cuda)

    def check():
        # fp64 reference in chunks of clouds (the grouped tensor is big)
        for b0 in range(0, B, 20):
            sl = slice(b0, min(B, b0 + 20)); nb = sl.stop - sl.start
            j = idx[sl].long()
            ar = torch.arange(nb, device=cuda).view(nb, 1, 1)
            hj = h.view(B, N, c1)[sl].double()[ar, j]
            d = xyz[sl].double()[ar, j] - centres[sl].double().unsqueeze(2)
            a1 = torch.relu(hj + d @ wxyz.double())
            a2 = torch.relu(a1 @ w2.double().t() + b2.double())
            a3 = torch.relu(a2 @ w3.double().t() + b3.double())
            ref = a3.max(dim=2)[0].view(nb * M, c3)
            got = out[b0 * M:(b0 + nb) * M, 4:4 + c3]
            assert rel_err(got, ref) < 5e-5, (b0, rel_err(got, ref))
        assert float(out[:, :4].max()) == -1.0 and float(out[:, 4 + c3:].max()) == -1.0   # neighbours untouched
    return l2, l3, (h, idx, xyz, centres, wxyz), out, check


@pytest.mark.parametrize("c1,c2,c3,ns,B,N,M", [(128, 128, 128, 64, 5, 512, 128), (16, 16, 32, 16, 2, 4096, 1024),
                                                (32, 32, 64, 32, 2, 4096, 1000), (64, 96, 128, 32, 3, 1024, 250),
                                                (128, 128, 128, 64, 300, 512, 128)])
def test_sa_fused_tc_vs_fp64(cuda, c1, c2, c3, ns, B, N, M):
    """whole SA scale on chip, row-major last layer (csrc/sa_fused_tc.cu): gather + layer-1 xyz half + layers
    2, 3 + max (shuffle-butterfly pooling)."""
    fz = load("fused")
    l2, l3, args, out, check = _sa_fused_case(cuda, fz, c1, c2, c3, ns, B, N, M)
    assert fz.sa_fused_supported(l2, l3, ns)
    saved, fz.SA_TRANSPOSED = fz.SA_TRANSPOSED, False
    try:
        fz.sa_fused_tc(*args, l2, l3, out[:, 4:4 + c3])
    finally:
        fz.SA_TRANSPOSED = saved
    check()


@pytest.mark.parametrize("c1,c2,c3,ns,B,N,M", [(128, 128, 128, 64, 5, 512, 128), (64, 96, 128, 32, 3, 1024, 250),
                                                (64, 64, 128, 16, 2, 1024, 999), (128, 128, 256, 64, 7, 128, 32),
                                                (128, 128, 256, 16, 3, 512, 50), (128, 128, 128, 128, 3, 512, 33),
                                                (128, 128, 256, 128, 2, 512, 17), (32, 16, 128, 32, 2, 256, 77),
                                                (16, 16, 32, 16, 2, 4096, 1024), (32, 32, 64, 32, 2, 4096, 1000),
                                                (128, 128, 256, 64, 400, 128, 32)])
def test_sa_fused_t_tc_vs_fp64(cuda, c1, c2, c3, ns, B, N, M):
    """whole SA scale on chip, TRANSPOSED last layer (csrc/sa_fused_t_tc.cu): W3 in tensor memory, in-thread
    max-pool; 128 output channels (double-buffered layer-2 accumulator) and 256 (two passes), every nsample,
    ragged last tiles."""
    fz = load("fused")
    l2, l3, args, out, check = _sa_fused_case(cuda, fz, c1, c2, c3, ns, B, N, M)
    saved_small, fz.SA_TRANSPOSED_SMALL = fz.SA_TRANSPOSED_SMALL, True       # < 128 channels: supported, not routed by default
    try:
        assert fz.SA_TRANSPOSED and fz.sa_fused_t_supported(l2, l3, ns)
        fz.sa_fused_tc(*args, l2, l3, out[:, 4:4 + c3])
    finally:
        fz.SA_TRANSPOSED_SMALL = saved_small
    check()
    if c3 == 128 and fz.sa_fused_supported(l2, l3, ns):      # the two kernels agree to rounding of the accumulation order
        out_t = out.clone()
        saved, fz.SA_TRANSPOSED = fz.SA_TRANSPOSED, False
        try:
            fz.sa_fused_tc(*args, l2, l3, out[:, 4:4 + c3])
        finally:
            fz.SA_TRANSPOSED = saved
        assert rel_err(out_t[:, 4:4 + c3], out[:, 4:4 + c3].double()) < 1e-6


@pytest.mark.parametrize("c1,c2,c3,ns,B,N,M", [(128, 128, 128, 64, 40, 512, 128), (128, 128, 256, 64, 30, 128, 32),
                                                (64, 96, 128, 32, 3, 1024, 250), (64, 64, 128, 16, 2, 1024, 999),
                                                (16, 16, 32, 16, 2, 2048, 777), (32, 32, 64, 32, 2, 2048, 512)])
def test_sa_fused_t_skips_padded_duplicates_exactly(cuda, c1, c2, c3, ns, B, N, M):
    """Duplicate-skipping mode (csrc/group_compact.cu + compact rows in sa_fused_t_tc.cu) on ball_query-shaped groups
    (cnt real neighbours, then copies of the first hit; some groups full, some with a single hit): the pooled output
    EQUALS the dense kernel's bit for bit (a max over a multiset does not see duplicates), and the fp64 reference."""
    fz = load("fused")
    l2, l3, args, out, check = _sa_fused_case(cuda, fz, c1, c2, c3, ns, B, N, M)
    h, idx, xyz, centres, wxyz = args
    g = torch.Generator(device="cpu").manual_seed(7)
    cnt = torch.randint(1, ns + 1, (B, M, 1), generator=g)
    cnt[0, :5] = ns
    cnt[-1, -5:] = 1
    k = torch.arange(ns).view(1, 1, ns)
    idx_cpu = idx.cpu()
    # distinct neighbours per group like ball_query: a random permutation prefix, padded with the first hit
    perm = torch.argsort(torch.rand((B, M, N), generator=g), dim=2)[:, :, :ns].to(torch.int32)
    idx_pad = torch.where(k < cnt, perm, perm[:, :, :1]).contiguous().to(cuda)
    args = (h, idx_pad, xyz, centres, wxyz)
    cm, jm, nrows = fz.group_compact(idx_pad, align=1)
    assert int(nrows.item()) == int(cnt.sum())
    u = int(nrows.item())
    assert torch.equal(cm[:u].cpu(), torch.repeat_interleave(torch.arange(B * M, dtype=torch.int32), cnt.view(-1)))
    assert torch.equal(jm[:u].cpu(), idx_pad.cpu()[k.expand(B, M, ns) < cnt])
    # align 8: every group topped up to a multiple of eight rows with copies of its first neighbour
    cm8, jm8, nrows8 = fz.group_compact(idx_pad, align=8)
    cnt8 = (cnt + 7) // 8 * 8
    u8 = int(nrows8.item())
    assert u8 == int(cnt8.sum())
    assert torch.equal(cm8[:u8].cpu(), torch.repeat_interleave(torch.arange(B * M, dtype=torch.int32), cnt8.view(-1)))
    k8 = k.expand(B, M, ns)
    want8 = torch.where(k8 < cnt, idx_pad.cpu(), idx_pad.cpu()[:, :, :1].expand(B, M, ns))[k8 < cnt8]
    assert torch.equal(jm8[:u8].cpu(), want8)
    # the two-launch compaction (prefix sum inside the list kernel) and the count / torch.cumsum / list path agree
    two = fz.COMPACT_TWO_LAUNCHES
    try:
        fz.COMPACT_TWO_LAUNCHES = not two
        cmo, jmo, nro = fz.group_compact(idx_pad, align=8)
    finally:
        fz.COMPACT_TWO_LAUNCHES = two
    assert int(nro.item()) == u8 and torch.equal(cmo[:u8], cm8[:u8]) and torch.equal(jmo[:u8], jm8[:u8])
    saved = (fz.SA_SKIP_DUPLICATES, fz.SA_SKIP_MIN_ROWS, fz.SA_TRANSPOSED_SMALL, fz.SA_COMPACT_ALIGN)
    try:
        fz.SA_SKIP_MIN_ROWS = 0
        fz.SA_TRANSPOSED_SMALL = True
        fz.SA_SKIP_DUPLICATES = False
        dense = torch.full((B * M, c3), -3.0, device=cuda)
        fz.sa_fused_tc(*args, l2, l3, dense)
        fz.SA_SKIP_DUPLICATES = True
        for align in (8, 1):              # group-wise and run-wise pooling epilogue
            fz.SA_COMPACT_ALIGN = align
            out.fill_(-1.0)
            fz.sa_fused_tc(*args, l2, l3, out[:, 4:4 + c3])
            assert torch.equal(out[:, 4:4 + c3], dense), "align %d" % align
            assert float(out[:, :4].max()) == -1.0 and float(out[:, 4 + c3:].max()) == -1.0
    finally:
        fz.SA_SKIP_DUPLICATES, fz.SA_SKIP_MIN_ROWS, fz.SA_TRANSPOSED_SMALL, fz.SA_COMPACT_ALIGN = saved


def test_linear_pre_two_layers_in_one_launch(cuda):
    """pn2_linear_pre_tc_f32: relu(relu(x[:, :5] W1^T + b1) W2^T + b2) with the 5-channel layer computed inside the
    producers, against fp64 and against the two separate launches; strided input rows (the pooled ROI tensor), ragged
    row counts."""
    fz = load("fused")
    g = torch.Generator(device="cpu").manual_seed(11)
    w1 = (torch.randn((128, 5), generator=g) / 5 ** 0.5).cuda(); b1 = torch.randn((128,), generator=g).cuda()
    w2 = (torch.randn((128, 128), generator=g) / 128 ** 0.5).cuda(); b2 = torch.randn((128,), generator=g).cuda()
    l1, l2 = fz.PackedLayer(w1, b1, True), fz.PackedLayer(w2, b2, True)
    for rows in (128 * 40, 1000, 77):
        wide = torch.randn((rows, 136), generator=g).cuda()             # rows of the padded pooled tensor
        got = fz.linear_pre(wide, 5, l1, l2)
        assert got is not None and got.shape == (rows, 128)
        x = wide[:, :5].double()
        ref = torch.relu(torch.relu(x @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double())
        assert rel_err(got, ref) < 3e-5
        two = fz.linear(fz.linear(wide[:, :5], l1), l2)
        assert rel_err(got, two.double()) < 3e-5
    # not the instantiated shape -> None (the caller runs the layers one by one)
    l1b = fz.PackedLayer((torch.randn((128, 6), generator=g)).cuda(), b1, True)
    assert fz.linear_pre(wide, 6, l1b, l2) is None


def test_rcnn_front_chain_in_one_launch(cuda):
    """pn2_rcnn_front_tc_f32: xyz_up [5 -> 128 -> 128] -> merge_down on cat[., rpn features] -> the per-point half of
    SA1's first layer, one launch, against fp64 and against the three separate launches; tile-multiple, ragged and tiny
    row counts (a CTA with a single tile, CTAs without any)."""
    fz = load("fused")
    g = torch.Generator(device="cpu").manual_seed(19)

    def mk(cout, cin, relu):
        w = (torch.randn((cout, cin), generator=g) / cin ** 0.5).cuda()
        b = torch.randn((cout,), generator=g).cuda()
        return w, b, fz.PackedLayer(w, b, relu)

    w1, b1, l1 = mk(128, 5, True)
    w2, b2, l2 = mk(128, 128, True)
    wm, bm, lm = mk(128, 256, True)
    ws, bs, ls = mk(128, 128, False)
    for rows in (128 * 148 * 3 + 128 * 5, 128 * 40, 1000, 77, 128):
        wide = torch.randn((rows, 136), generator=g).cuda()              # [x y z mask depth | pad | 128 features]
        wide[:, 5:8] = 0
        assert fz.rcnn_front_supported(wide, 5, 8, l1, l2, lm, ls)
        got = fz.rcnn_front(wide, 8, l1, l2, lm, ls)
        x = wide.double()
        a = torch.relu(torch.relu(x[:, :5] @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double())
        m = torch.relu(torch.cat((a, x[:, 8:]), dim=1) @ wm.double().t() + bm.double())
        ref = m @ ws.double().t() + bs.double()
        e = rel_err(got, ref)
        assert e < 1e-4, (rows, e)
        assert e < 5e-5, "three BF16x3 layers should stay near 1e-5: %g at %d rows" % (e, rows)
        three = fz.linear(fz.linear_cat(fz.linear_pre(wide, 5, l1, l2), wide[:, 8:], lm), ls)
        assert rel_err(got, three.double()) < 5e-5
    # pad columns hold garbage (the one-launch pooling kernel writes zeros, older callers may not): never read
    wide[:, 5:8] = float("nan")
    assert bool(torch.isfinite(fz.rcnn_front(wide, 8, l1, l2, lm, ls)).all())
    # not the instantiated shape
    assert not fz.rcnn_front_supported(wide, 5, 8, l1, l2, lm, fz.PackedLayer(ws, bs, True))
