"""csrc/glue.cu + pn2_roipool3d_canon_f32: every one-launch stage is BIT-IDENTICAL to the torch composition it replaces
(bbox_transform.decode_bbox_target_torch, ProposalLayer's torch flow, RCNNNet._pool_rois_padded, Detector.postprocess_torch),
which are themselves the reference's statements (tests/test_host_utils_cpu.py, test_refnet_vs_port_cpu.py).
The kernels must reproduce IEEE single precision exactly: decoded boxes feed thresholds and NMS."""
import numpy as np
import pytest
import torch

from conftest import load

pytestmark = pytest.mark.gpu
synthetic = load("synthetic")


@pytest.fixture(scope="module")
def model(cuda):
    return load("inference").build_model(seed=0, device=cuda)


def _bits_equal(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = a.view(torch.int32) == b.view(torch.int32)
    # +0.0 and -0.0 compare equal as floats; everything else must agree in every bit
    same |= (a == 0) & (b == 0)
    assert bool(same.all()), "%s: %d of %d elements differ (max |diff| %.3e)" % (
        what, int((~same).sum()), same.numel(), float((a - b).abs().max()))


def _reg_like(rows, c, seed, cuda, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    reg = torch.randn((rows, c), generator=g) * scale
    # ties inside an argmax window and a NaN-free but extreme row
    reg[3, 0:4] = 1.25
    reg[7, :] = 0.0
    return reg.to(cuda)


@pytest.mark.parametrize("fine,ybin,ryfine", [(True, False, False), (False, False, False), (True, True, True), (True, False, True)])
def test_decode_bbox_matches_torch_bitwise(cuda, fine, ybin, ryfine):
    bt, glue, cfg = load("bbox_transform"), load("glue"), load("config").cfg
    rows = 5000
    nb = int(3.0 / 0.5) * 2
    c = nb * (4 if fine else 2) + (2 * int(0.5 / 0.25) * 2 if ybin else 1) + 2 * 12 + 3
    reg = _reg_like(rows, c, 3, cuda)
    g = torch.Generator(device="cpu").manual_seed(9)
    anchor = torch.from_numpy(cfg.CLS_MEAN_SIZE[0]).to(cuda)
    for roi_dim in (3, 7):
        roi = (torch.randn((rows, roi_dim), generator=g) * torch.tensor([20.0, 1.0, 30.0, 0.3, 0.3, 0.5, 2.0][:roi_dim])).to(cuda)
        want = bt.decode_bbox_target_torch(roi, reg, 3.0, 0.5, 12, anchor, get_xz_fine=fine, get_y_by_bin=ybin, loc_y_scope=0.5,
                                     loc_y_bin_size=0.25, get_ry_fine=ryfine)
        got = glue.decode_bbox(roi, reg, 3.0, 0.5, 12, cfg.CLS_MEAN_SIZE[0], get_xz_fine=fine, get_y_by_bin=ybin,
                               loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=ryfine)
        if roi_dim == 7 and not torch.equal(got, want):
            # report which rounding of the K = 2 matmul torch used on this box before failing
            agree = {}
            for mode in (0, 1, 2):
                glue.ROT_MODE, keep = mode, glue.ROT_MODE
                agree[mode] = float((glue.decode_bbox(roi, reg, 3.0, 0.5, 12, cfg.CLS_MEAN_SIZE[0], get_xz_fine=fine,
                                                      get_y_by_bin=ybin, loc_y_scope=0.5, loc_y_bin_size=0.25,
                                                      get_ry_fine=ryfine) == want).float().mean())
                glue.ROT_MODE = keep
            print("fraction of equal elements per rot_mode:", agree)
        _bits_equal(got, want, "decode roi_dim=%d" % roi_dim)
    # the proposal layer's variant: y moved to the bottom face
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 1, rows, seed=2)[0]).to(cuda)
    want = bt.decode_bbox_target_torch(xyz, reg, 3.0, 0.5, 12, anchor, get_xz_fine=fine, get_y_by_bin=ybin, get_ry_fine=ryfine)
    want[:, 1] += want[:, 3] / 2
    got = glue.decode_bbox(xyz, reg, 3.0, 0.5, 12, cfg.CLS_MEAN_SIZE[0], get_xz_fine=fine, get_y_by_bin=ybin,
                           get_ry_fine=ryfine, y_bottom=True)
    _bits_equal(got, want, "decode y_bottom")


@pytest.mark.parametrize("case", ["random", "no_far_band", "few_near", "all_out_of_range"])
def test_proposal_layer_kernels_match_torch_flow(cuda, model, case):
    """ProposalLayer._forward_kernels == _forward_batched (torch, batched) == the reference's per-scene flow."""
    glue = load("glue")
    B, N = 3, 16384
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", B, N, seed=21)).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(17)
    scores = torch.randn((B, N), generator=g).to(cuda)
    reg = (torch.randn((B, N, 76), generator=g) * 0.5).to(cuda)
    if case == "no_far_band":
        xyz[..., 2] = xyz[..., 2].clamp(max=30.0)         # nothing decodes beyond 40 m: the far band borrows
    elif case == "few_near":
        xyz[1, :, 2] = xyz[1, :, 2].abs() + 45.0          # scene 1: everything in the far band
        xyz[2, 100:, 2] = -5.0                             # scene 2: 100 candidates in total
    elif case == "all_out_of_range":
        xyz[0, :, 2] = 200.0
    scores[0, 10:20] = 0.75                                # tied scores: the sort's order is shared by both paths
    pl = model.rpn.proposal_layer
    old = glue.ENABLED
    try:
        glue.ENABLED = True
        rois, roi_scores = pl(scores, reg, xyz)
        glue.ENABLED = False
        rois_t, roi_scores_t = pl(scores, reg, xyz)
        # the reference's per-scene flow asserts on a scene whose near band is empty (proposal_layer.py:94)
        if case in ("random", "no_far_band"):
            pl.fused = False
            rois_r, roi_scores_r = pl(scores, reg, xyz)
        else:
            rois_r, roi_scores_r = rois_t, roi_scores_t
    finally:
        glue.ENABLED = old
        pl.fused = True
    assert rois.shape == (B, 100, 7)
    _bits_equal(rois, rois_t, "rois vs batched torch (%s)" % case)
    _bits_equal(roi_scores, roi_scores_t, "roi scores vs batched torch")
    _bits_equal(rois, rois_r, "rois vs per-scene reference flow (%s)" % case)
    _bits_equal(roi_scores, roi_scores_r, "roi scores vs per-scene reference flow")


def test_rcnn_input_stage_one_launch_matches_torch_flow(cuda, model):
    """pn2_roipool3d_canon_f32 == enlarge_box3d + pn2_roipool3d_split_f32 + the torch canonical transform, incl. empty ROIs."""
    glue, cfg = load("glue"), load("config").cfg
    B, N, C = 2, 16384, 128
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", B, N, seed=5)).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(23)
    feats = torch.randn((B, N, C), generator=g).to(cuda)
    scores = torch.randn((B, N), generator=g).to(cuda) * 2
    # ROIs centred on cloud points (non-empty), a few far away (empty)
    pick = torch.randint(0, N, (B, 100), generator=g).to(cuda)
    centres = torch.gather(xyz, 1, pick.unsqueeze(-1).expand(-1, -1, 3))
    rois = torch.cat((centres, torch.tensor([1.5, 1.6, 3.9], device=cuda).expand(B, 100, 3),
                      (torch.rand((B, 100, 1), generator=g) * 6.28 - 3.14).to(cuda)), dim=2).contiguous()
    rois[:, 1, 0:3] = torch.tensor([500.0, 0.0, 500.0], device=cuda)
    rois[1, 7, 0:3] = torch.tensor([-300.0, 2.0, 80.0], device=cuda)
    seg = (torch.sigmoid(scores) > cfg.RPN.SCORE_THRESH).float()
    depth = torch.norm(xyz, p=2, dim=2)
    data = {'rpn_xyz': xyz, 'rpn_features': feats, 'seg_mask': seg, 'roi_boxes3d': rois, 'pts_depth': depth,
            'rpn_scores_raw': scores, 'seg_thresh': cfg.RPN.SCORE_THRESH}
    net = model.rcnn_net
    want = net._pool_rois_padded(data)
    got = net._pool_rois_canonical(data)
    agree = {}
    if not torch.equal(got[..., 0:3], want[..., 0:3]):
        keep = glue.ROT_MODE_POOL
        for mode in (0, 1, 2):
            glue.ROT_MODE_POOL = mode
            agree[mode] = float((net._pool_rois_canonical(data)[..., 0:3] == want[..., 0:3]).float().mean())
        glue.ROT_MODE_POOL = keep
        print("canonical xyz: fraction equal per rot_mode", agree)
    _bits_equal(got[..., 3:], want[..., 3:], "pooled features")
    _bits_equal(got[..., 0:3], want[..., 0:3], "canonical xyz")
    assert float(got[1].abs()[:, 3:].sum()) == 0.0          # empty ROI: zero features, transformed zero point


def test_postprocess_kernels_match_torch(cuda, model):
    inf, glue = load("inference"), load("glue")
    det = inf.Detector(model, cuda, use_graph=False, depth=1)
    B, M = 4, 100
    g = torch.Generator(device="cpu").manual_seed(31)
    centres = torch.rand((B, 12, 3), generator=g) * torch.tensor([40.0, 1.0, 60.0]) + torch.tensor([-20.0, 1.0, 5.0])
    idx = torch.randint(0, 12, (B, M), generator=g)
    rois = torch.cat((torch.gather(centres, 1, idx.unsqueeze(-1).expand(-1, -1, 3)) + torch.randn((B, M, 3), generator=g) * 0.3,
                      torch.tensor([1.5, 1.6, 3.9]).expand(B, M, 3) + torch.randn((B, M, 3), generator=g) * 0.1,
                      torch.rand((B, M, 1), generator=g) * 6.28 - 3.14), dim=2).to(cuda).contiguous()
    rois[0, 90:] = 0.0                                       # zero-padded proposals
    cls = (torch.randn((B * M, 1), generator=g) * 2).to(cuda)
    cls[5:9] = 1.5                                           # tied scores: stable order
    cls[M:2 * M] = -20.0                                     # scene 1: nothing above the threshold
    reg = _reg_like(B * M, 46, 41, cuda, scale=0.5)
    ret = {'rois': rois, 'rcnn_cls': cls, 'rcnn_reg': reg}
    rec_t, num_t = det.postprocess_torch(ret, B)
    old = glue.ENABLED
    try:
        glue.ENABLED = True
        rec, num = det.postprocess(ret, B)
    finally:
        glue.ENABLED = old
    assert torch.equal(num.cpu(), num_t.cpu()), (num, num_t)
    assert int(num[1]) == 0 and int(num.sum()) > 10
    _bits_equal(rec, rec_t, "detection records")


def test_detector_with_and_without_glue_kernels_is_identical(cuda, model):
    """whole step: the one-launch stages change no bit of the detections."""
    inf, glue = load("inference"), load("glue")
    pts = torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=77)).to(cuda)
    det = inf.Detector(model, cuda, use_graph=False, depth=1)
    old = glue.ENABLED
    try:
        glue.ENABLED = False
        rec0, num0 = det.detect_device(pts)
        glue.ENABLED = True
        rec1, num1 = det.detect_device(pts)
    finally:
        glue.ENABLED = old
    assert torch.equal(num0, num1)
    _bits_equal(rec1, rec0, "records with / without the glue kernels")


def test_whole_forward_graph_replay_equals_eager(cuda):
    """PointRCNN.graph_forward (what the drop-in tree of the unmodified eval_rcnn.py switches on): replaying the captured
    forward gives the eager launches' outputs bit for bit, for new data in the same shape, for a second shape, and again
    after the weights changed (load_state_dict drops the captured graphs)."""
    inf = load("inference")
    model = inf.build_model(seed=0, device=cuda)
    keys = ("rpn_cls", "rpn_reg", "backbone_features", "rois", "roi_scores_raw", "seg_result", "rcnn_cls", "rcnn_reg")

    def both(pts):
        with torch.no_grad():
            model.graph_forward = False
            want = model({"pts_input": pts})
            model.graph_forward = True
            got = model({"pts_input": pts})
            again = model({"pts_input": pts})
        model.graph_forward = False
        for k in keys:
            assert torch.equal(got[k], want[k]), k
            assert torch.equal(again[k], want[k]) and again[k].data_ptr() != got[k].data_ptr(), k
        return got

    a = both(torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=5)).to(cuda))
    b = both(torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=6)).to(cuda))      # same shape: a replay
    assert not torch.equal(a["rois"], b["rois"])
    both(torch.from_numpy(synthetic.make_clouds("lidar", 3, 16384, seed=7)).to(cuda))           # another shape
    assert len(model._graphs) == 2
    state = {k: v.clone() for k, v in model.state_dict().items()}
    state["rcnn_net.cls_layer.2.conv.bias"] += 1.0
    model.load_state_dict(state)
    assert len(model._graphs) == 0
    c = both(torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=5)).to(cuda))
    assert not torch.equal(c["rcnn_cls"], a["rcnn_cls"])


@pytest.mark.parametrize("b,n", [(16, 16384), (3, 5000), (2, 1), (1, 2), (4, 777)])
def test_argsort_desc_is_torch_sort(cuda, b, n):
    """pn2_argsort_desc_f32 == torch.sort(scores, dim=1, descending=True)[1] (proposal_layer.py:26), with many exactly
    equal scores (ties go to the lower index, as the stable radix sort of torch leaves them)."""
    glue = load("glue")
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + n)
    scores = torch.randn((b, n), generator=g)
    scores[:, ::3] = torch.round(scores[:, ::3] * 2) / 2            # a third of the scores on a 0.5 lattice: lots of ties
    scores = scores.to(cuda)
    want = torch.sort(scores, dim=1, descending=True, stable=True)[1]
    got = glue.argsort_desc(scores)
    assert got.dtype == torch.int64
    assert torch.equal(torch.gather(scores, 1, got), torch.gather(scores, 1, want))       # a descending order of the scores
    assert torch.equal(torch.sort(got, dim=1)[0], torch.arange(n, device=cuda).expand(b, n))   # a permutation
    assert torch.equal(got, want), "ties: %d positions differ from the stable order" % int((got != want).sum())
    assert torch.equal(got, torch.sort(scores, dim=1, descending=True)[1])      # the call the reference makes


@pytest.mark.parametrize("rotated", [False, True])
def test_nms_pair_launch_equals_two_launches(cuda, rotated):
    """pn2_nms_bev_pair_f32 (both distance bands of the proposal layer in one launch) == two pn2_nms_bev_f32 calls."""
    glue = load("glue")
    g = torch.Generator(device="cpu").manual_seed(11)

    def problems(p, n):
        c = torch.rand((p, n, 2), generator=g) * 20.0
        wl = torch.rand((p, n, 2), generator=g) * 3.0 + 0.5
        ry = (torch.rand((p, n, 1), generator=g) - 0.5) * 6.0
        bev = torch.cat((c - wl / 2, c + wl / 2, ry), dim=2).contiguous().to(cuda)        # x1, y1, x2, y2, ry
        cnt = torch.randint(n // 2, n + 1, (p,), generator=g, dtype=torch.int32).to(cuda)
        return bev, cnt

    bev0, cnt0 = problems(5, 300 if rotated else 900)
    bev1, cnt1 = problems(3, 100 if rotated else 400)
    k0, n0 = glue.nms_raw(bev0, cnt0, 0.4, rotated, 40)
    k1, n1 = glue.nms_raw(bev1, cnt1, 0.4, rotated, 17)
    pk0, pn0, pk1, pn1 = glue.nms_raw_pair(bev0, cnt0, 40, bev1, cnt1, 17, 0.4, rotated)
    assert torch.equal(pn0, n0) and torch.equal(pn1, n1)
    for k, pk, n in ((k0, pk0, n0), (k1, pk1, n1)):
        for i in range(k.shape[0]):
            assert torch.equal(pk[i, :int(n[i])], k[i, :int(n[i])])
