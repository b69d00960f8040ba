"""GPU parity of the shared-MLP kernels and of the fused SA / FP modules and the whole
PointRCNN forward against the plain fp32 PyTorch composition the reference uses
(group -> 1x1 conv [+BN] + ReLU -> max_pool2d; TF32 disabled).  Tolerance for feature
tensors: 1e-4 relative (BASELINE.json north_star), written as
    |a - b| <= 1e-4 * |b| + 1e-5 * max|b|
Index-valued intermediates (FPS / ball query) are shared by both paths, so the comparison is
stage-wise with teacher forcing where a discrete decision (argmax bin, NMS) sits in between."""
import numpy as np
import pytest
import torch

from conftest import load

pytestmark = pytest.mark.gpu
synthetic = load("synthetic")


@pytest.fixture(autouse=True, params=["tc", "ffma"])
def _engine_and_fp32_reference(request):
    """every test runs twice: shared-MLP layers on the tcgen05 tensor cores (BF16x3 split, the
    default) and on the exact-fp32 CUDA-core kernels; the PyTorch side has TF32 disabled."""
    fz = load("fused")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, fz.MLP_ENGINE)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fz.set_mlp_engine(request.param)
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old[:2]
    fz.set_mlp_engine(old[2])


def assert_feat_close(a, b, what=""):
    """1e-4 relative.  Exact-fp32 engine: element-wise, |a-b| <= 1e-4 |b| + 1e-5 max|b| (only the
    summation order differs).  Tensor-core engine: relative to the tensor scale,
    |a-b| <= 1e-4 max|b| (split-bf16 products carry ~1e-5 of the scale per layer)."""
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = float(b.abs().max())
    err = (a - b).abs()
    if load("fused").MLP_ENGINE == "ffma":
        bound = 1e-4 * b.abs() + 1e-5 * scale
    else:
        bound = torch.full_like(err, 1e-4 * scale)
    bad = err > bound
    print("%s [%s]: max err / scale = %.2e" % (what, load("fused").MLP_ENGINE, float(err.max()) / max(scale, 1e-30)))
    assert not bool(bad.any()), "%s: %d / %d outside 1e-4 rel, max err %.3e at scale %.3e" % (
        what, int(bad.sum()), bad.numel(), float(err.max()), scale)


@pytest.mark.parametrize("rows,cin,cout,relu", [(1000, 3, 16, True), (4096, 131, 128, True), (777, 256, 46, False),
                                                (128, 1536, 512, True), (5, 7, 1, False), (2048, 96, 130, True)])
def test_linear_vs_torch(cuda, rows, cin, cout, relu):
    fz = load("fused")
    g = torch.Generator(device="cpu").manual_seed(rows + cin)
    x = torch.randn((rows, cin), generator=g).to(cuda)
    w = (torch.randn((cout, cin), generator=g) / cin ** 0.5).to(cuda)
    b = torch.randn((cout,), generator=g).to(cuda)
    layer = fz.PackedLayer(w, b, relu)
    y = fz.linear(x, layer)
    ref = (x.double() @ w.double().t() + b.double())
    ref = (ref.clamp_min(0) if relu else ref).float()
    assert_feat_close(y, ref, "linear")


def test_linear_strided_views_residual_and_pool(cuda):
    fz = load("fused")
    g = torch.Generator(device="cpu").manual_seed(0)
    big = torch.randn((2048, 200), generator=g).to(cuda)
    w = (torch.randn((64, 100), generator=g) / 10).to(cuda)
    b = torch.randn((64,), generator=g).to(cuda)
    res = torch.randn((2048, 64), generator=g).to(cuda)
    layer = fz.PackedLayer(w, b, True)
    x = big[:, 50:150]                                    # column slice: ld 200, not 16-byte aligned
    out = torch.zeros((2048, 96), device=cuda)
    fz.linear(x, layer, out=out[:, 32:], res=res)         # written into a column slice
    ref = torch.relu(x.double() @ w.double().t() + b.double() + res.double()).float()
    assert_feat_close(out[:, 32:], ref, "residual")
    assert float(out[:, :32].abs().max()) == 0.0
    for pool in (16, 32, 64, 128):
        y = fz.linear(x, layer, pool=pool)
        ref = torch.relu(x.double() @ w.double().t() + b.double()).float().view(2048 // pool, pool, 64).max(1)[0]
        assert_feat_close(y, ref, "pool %d" % pool)


def _sa_msg(cuda, cin, bn, npoint=256, radii=(0.8, 1.6), nsamples=(16, 32), mlps=((16, 16, 32), (32, 32, 64))):
    mods = load("pointnet2_modules")
    torch.manual_seed(0)
    sa = mods.PointnetSAModuleMSG(npoint=npoint, radii=list(radii), nsamples=list(nsamples),
                                  mlps=[[cin] + list(m) for m in mlps], bn=bn, use_xyz=True).to(cuda)
    if bn:  # non-trivial running statistics so that the fold is exercised
        for m in sa.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    return sa.eval()


@pytest.mark.parametrize("cin,bn", [(0, True), (5, True), (64, False)])
def test_sa_msg_fused_vs_reference_composition(cuda, cin, bn):
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 3, 2048, seed=1024)).to(cuda)
    feats = torch.randn((3, cin, 2048), device=cuda) if cin else None
    sa = _sa_msg(cuda, cin, bn)
    with torch.no_grad():
        new_xyz, out = sa(xyz, feats)
        sa.fused = False
        ref_xyz, ref = sa(xyz, feats)
    assert torch.equal(new_xyz, ref_xyz)
    assert out.shape == (3, 96, 256)
    assert_feat_close(out, ref, "SA-MSG")


def test_sa_single_scale_and_group_all(cuda):
    mods = load("pointnet2_modules")
    torch.manual_seed(1)
    xyz = (torch.rand((40, 512, 3), device=cuda) - 0.5) * 2
    feats = torch.randn((40, 128, 512), device=cuda)
    sa1 = mods.PointnetSAModule(npoint=128, radius=0.2, nsample=64, mlp=[128, 128, 128, 128], bn=False).to(cuda).eval()
    sa3 = mods.PointnetSAModule(mlp=[128, 256, 256, 512], bn=False).to(cuda).eval()
    with torch.no_grad():
        x1, f1 = sa1(xyz, feats)
        x3, f3 = sa3(x1[:, :64].contiguous(), f1[:, :, :64].contiguous())
        sa1.fused = sa3.fused = False
        rx1, rf1 = sa1(xyz, feats)
        rx3, rf3 = sa3(x1[:, :64].contiguous(), f1[:, :, :64].contiguous())
    assert torch.equal(x1, rx1) and x3 is None and rx3 is None
    assert_feat_close(f1, rf1, "SA ns=64")
    assert f3.shape == (40, 512, 1)
    assert_feat_close(f3, rf3, "GroupAll")


def test_fp_fused_vs_reference_composition(cuda):
    mods = load("pointnet2_modules")
    torch.manual_seed(2)
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 2, 4096, seed=5)).to(cuda)
    known = xyz[:, ::4].contiguous()
    for c1 in (0, 96):
        fp = mods.PointnetFPModule(mlp=[256 + c1, 128, 128], bn=True).to(cuda).eval()
        uf = torch.randn((2, c1, 4096), device=cuda) if c1 else None
        kf = torch.randn((2, 256, 1024), device=cuda)
        with torch.no_grad():
            out = fp(xyz, known, uf, kf)
            fp.fused = False
            ref = fp(xyz, known, uf, kf)
        assert_feat_close(out, ref, "FP c1=%d" % c1)


@pytest.fixture(scope="module")
def model(cuda):
    load("config").use_default_yaml("rcnn")
    torch.manual_seed(0)
    net = load("net.point_rcnn").PointRCNN(num_classes=2, use_xyz=True, mode="TEST").to(cuda).eval()
    # random-init BN has identity statistics; perturb so folding is tested on the real graph
    g = torch.Generator(device="cpu").manual_seed(3)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return net


def _set_fused(net, flag):
    for m in net.modules():
        if hasattr(m, "fused"):
            m.fused = flag


def test_rpn_stage_fused_vs_reference_composition(cuda, model):
    pts = torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=1024)).to(cuda)
    with torch.no_grad():
        _set_fused(model, True)
        a = model.rpn({"pts_input": pts})
        _set_fused(model, False)
        b = model.rpn({"pts_input": pts})
        _set_fused(model, True)
    assert torch.equal(a["backbone_xyz"], b["backbone_xyz"])
    assert a["rpn_cls"].shape == (2, 16384, 1) and a["rpn_reg"].shape == (2, 16384, 76)
    assert a["backbone_features"].shape == (2, 128, 16384)
    for k in ("backbone_features", "rpn_cls", "rpn_reg"):
        assert_feat_close(a[k], b[k], k)


def test_rcnn_stage_fused_vs_reference_composition(cuda, model):
    """teacher-forced: both paths get the same RPN outputs and ROIs."""
    B, N = 2, 16384
    xyz_h = synthetic.make_clouds("lidar", B, N, seed=666)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    rng = np.random.RandomState(0)
    rois = np.zeros((B, 100, 7), np.float32)
    for b in range(B):
        for m in range(90):  # the last 10 stay zero boxes, as the proposal layer pads them
            p = xyz_h[b, rng.randint(0, N)]
            rois[b, m] = [p[0], p[1] + 0.8, p[2], 1.5, 1.6, 3.9, rng.uniform(-np.pi, np.pi)]
    info = {"rpn_xyz": xyz, "rpn_features": torch.randn((B, N, 128), device=cuda),
            "seg_mask": (torch.rand((B, N), device=cuda) > 0.5).float(), "roi_boxes3d": torch.from_numpy(rois).to(cuda),
            "pts_depth": torch.norm(xyz, p=2, dim=2)}
    with torch.no_grad():
        _set_fused(model, True)
        a = model.rcnn_net(dict(info))
        _set_fused(model, False)
        b = model.rcnn_net(dict(info))
        _set_fused(model, True)
    assert a["rcnn_cls"].shape == (B * 100, 1) and a["rcnn_reg"].shape == (B * 100, 46)
    assert_feat_close(a["rcnn_cls"], b["rcnn_cls"], "rcnn_cls")
    assert_feat_close(a["rcnn_reg"], b["rcnn_reg"], "rcnn_reg")


def test_full_forward_runs_and_is_deterministic(cuda, model):
    pts = torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=7)).to(cuda)
    with torch.no_grad():
        o1 = model({"pts_input": pts})
        o2 = model({"pts_input": pts})
    for k in ("rpn_cls", "rpn_reg", "backbone_xyz", "backbone_features", "rois", "roi_scores_raw", "seg_result",
              "rcnn_cls", "rcnn_reg"):
        assert k in o1, k
        assert torch.equal(o1[k], o2[k]), k
    assert o1["rois"].shape == (2, 100, 7) and o1["rcnn_reg"].shape == (200, 46)
    assert torch.isfinite(o1["rcnn_reg"]).all()


def test_detector_pipelined_submit_collect_matches_single_stream(cuda, model):
    """Detector.submit/collect (two batches in flight, one stream + CUDA graph per slot) returns, batch for
    batch, exactly the records of the one-at-a-time detect()."""
    inf = load("inference")
    batches = [torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=20 + i)).pin_memory() for i in range(5)]
    ref = inf.Detector(model, cuda, use_graph=False, depth=1)
    want = []
    for b in batches:
        rec, num = ref.detect(b)
        want.append((rec.clone(), num.clone()))
    for use_graph in (False, True):
        det = inf.Detector(model, cuda, use_graph=use_graph, depth=2)
        got = [(rec.clone(), num.clone()) for rec, num in det.detect_stream(batches)]
        assert len(got) == len(want)
        for i, ((r0, n0), (r1, n1)) in enumerate(zip(want, got)):
            assert torch.equal(n0, n1), (use_graph, i)
            assert torch.equal(r0, r1), (use_graph, i)
        # device-side collect: the current stream waits, views of the slot's buffers
        t = det.submit(batches[0].to(cuda))
        rec, num = det.collect(t, host=False)
        assert torch.equal(rec.cpu(), want[0][0]) and torch.equal(num.cpu(), want[0][1])
        det.drain()


def test_proposal_layer_vs_reference_nms_composition(cuda, model, legacy):
    """ProposalLayer with the device NMS == the reference recipe (sort, band split, top-k,
    legacy mask kernel + host greedy, first 70/30) on the same scores / regression."""
    cfg = load("config").cfg
    ku = load("kitti_utils")
    bt = load("bbox_transform")
    B, N = 2, 16384
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", B, N, seed=11)).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(5)
    scores = torch.randn((B, N), generator=g).to(cuda)
    reg = torch.randn((B, N, 76), generator=g).to(cuda)
    pl = model.rpn.proposal_layer
    rois, roi_scores = pl(scores, reg, xyz)
    assert rois.shape == (B, 100, 7)
    props = bt.decode_bbox_target_torch(xyz.view(-1, 3), reg.view(-1, 76), anchor_size=pl.MEAN_SIZE, loc_scope=cfg.RPN.LOC_SCOPE,
                                  loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                                  get_xz_fine=True, get_y_by_bin=False, get_ry_fine=False)
    props[:, 1] += props[:, 3] / 2
    props = props.view(B, N, 7)
    for k in range(B):
        order = torch.sort(scores[k], descending=True)[1]
        so, po = scores[k][order], props[k][order]
        dist = po[:, 2]
        outs, outp = [], []
        for lo, hi, pre, post in ((0, 40.0, 6300, 70), (40.0, 80.0, 2700, 30)):
            band = (dist > lo) & (dist <= hi)
            cs, cp = so[band][:pre], po[band][:pre]
            bev = ku.boxes3d_to_bev_torch(cp).contiguous()
            o2 = cs.sort(0, descending=True)[1]
            keep = legacy.greedy_from_mask(legacy.nms_mask(bev[o2].contiguous(), cfg.TEST.RPN_NMS_THRESH, normal=True).cpu(), len(o2))
            keep = o2[torch.from_numpy(keep).to(cuda)][:post]
            outs.append(cs[keep]); outp.append(cp[keep])
        rs, rp = torch.cat(outs), torch.cat(outp)
        assert torch.equal(rois[k, :len(rp)], rp) and torch.equal(roi_scores[k, :len(rs)], rs)
        assert float(rois[k, len(rp):].abs().sum()) == 0.0
