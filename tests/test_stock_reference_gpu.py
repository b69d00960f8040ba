"""The reference's OWN Python on a GPU (oracle/refnet_gpu.py), live on the box, against the product:

  * backend "legacy" -- unmodified lib/net/*.py, pointnet2_lib/pointnet2/*.py, lib/rpn/proposal_layer.py over the
    reference's unmodified kernels (oracle/_ref/libpn2_legacy.so): the bench's `--impl reference` arm.  The fused
    sm_100a path must agree with it stage by stage within 1e-4 of the tensor scale (RCNN teacher-forced with the
    reference's ROIs and mask, as in test_refnet_golden_gpu.py), at the fixture size AND at the benchmarked
    configuration B=16 x 16384 points, where the detections of the CUDA-graph / batches-in-flight Detector are also
    matched box by box against the reference's eval loop.
  * backend "b200"   -- the same unmodified Python with `pointnet2_cuda` / `iou3d_cuda` / `roipool3d_cuda` bound to the
    package's ctypes stubs: every index-valued tensor must equal the legacy run bit for
    bit given the same inputs, i.e. the reference runs unchanged on the new kernels.
cuDNN / cuBLAS TF32 is switched off for the reference side in these tests (the bench leaves torch's defaults)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_refnet_fixture as fx             # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refnet(cuda):
    from oracle import refnet_gpu
    if not refnet_gpu.available("legacy"):
        pytest.skip("stock reference tree (baseline/_ref/pointrcnn) or libpn2_legacy.so not available")
    return refnet_gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _close(got, want, rel, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    print("%s: max err / scale = %.2e" % (what, err / scale))
    assert err <= rel * scale, "%s: max err %.3e at scale %.3e" % (what, err, scale)


def _stagewise(model, ref_out, pts):
    """fused path vs the reference's forward outputs on the same clouds (see test_refnet_golden_gpu.py)."""
    with torch.no_grad():
        rpn = model.rpn({"pts_input": pts})
        _close(rpn["rpn_cls"], ref_out["rpn_cls"], 1e-4, "rpn_cls")
        _close(rpn["rpn_reg"], ref_out["rpn_reg"], 1e-4, "rpn_reg")
        _close(rpn["backbone_features"], ref_out["backbone_features"], 1e-4, "backbone_features")
        scores = rpn["rpn_cls"][:, :, 0]
        seg = (torch.sigmoid(scores) > 0.3).float()
        gold_seg = ref_out["seg_result"].float()
        flips = seg != gold_seg
        thr = float(np.log(0.3 / 0.7))
        assert bool(((scores - thr).abs()[flips] < 1e-4 * float(ref_out["rpn_cls"].abs().max())).all())
        assert int(flips.sum()) <= max(4, pts.shape[0] * pts.shape[1] // 4000)
        out = model.rcnn_net({"rpn_xyz": rpn["backbone_xyz"], "rpn_features": rpn["backbone_features"].permute(0, 2, 1),
                              "seg_mask": gold_seg, "roi_boxes3d": ref_out["rois"],
                              "pts_depth": torch.norm(rpn["backbone_xyz"], p=2, dim=2)})
        _close(out["rcnn_cls"], ref_out["rcnn_cls"], 1e-4, "rcnn_cls")
        _close(out["rcnn_reg"], ref_out["rcnn_reg"], 1e-4, "rcnn_reg")


def test_stock_reference_on_legacy_kernels_vs_fused_path(cuda, refnet):
    model = fx.seeded_model(cuda)
    ref = refnet.Reference(model.state_dict(), cuda, backend="legacy")
    pts = fx.scenes().to(cuda)
    ref_out = ref.forward(pts)
    # the live GPU run of the reference reproduces the committed CPU golden of the same network (index tensors are
    # bit-exact between the reference kernels and their C restatements; cuDNN fp32 vs torch CPU convs differ in
    # summation order only)
    z = np.load(os.path.join(ROOT, "tests", "golden", "refnet_forward.npz"))
    _close(ref_out["rpn_cls"], torch.from_numpy(z["rpn_cls"]).to(cuda), 2e-5, "reference GPU vs CPU golden rpn_cls")
    _stagewise(model, ref_out, pts)


def test_benchmarked_configuration_vs_stock_reference(cuda, refnet):
    """BASELINE configs[3] as bench.py times it: B=16 x 16384, CUDA graph, 3 batches in flight."""
    syn, inf = load("synthetic"), load("inference")
    model = inf.build_model(seed=0, device=cuda)
    ref = refnet.Reference(model.state_dict(), cuda, backend="legacy")
    host = torch.from_numpy(syn.make_clouds("lidar", 16, 16384, seed=1024)).pin_memory()
    pts = host.to(cuda)
    ref_out = ref.forward(pts)
    _stagewise(model, ref_out, pts)
    ref_dets = ref.eval_batch(host)
    det = inf.Detector(model, cuda, use_graph=True, depth=3)
    tickets = [det.submit(host, to_host=True) for _ in range(3)]
    got = None
    for t in tickets:
        h_rec, h_cnt = det.collect(t)
        cur = inf.records_to_lists(h_rec.clone(), h_cnt.clone())
        if got is not None:      # every slot's graph gives the same answer
            assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(got, cur))
        got = cur
    sys.path.insert(0, ROOT)
    import bench
    m = bench.match_detections(ref_dets, got, tol=2e-3)
    print("detections at B=16 x 16384 vs the stock reference:", m)
    assert m["total"] > 0
    assert m["matched"] >= 0.95 * m["total"] and m["extra"] <= 0.05 * m["total"] + 1


def test_reference_python_runs_unchanged_on_the_b200_stubs(cuda, refnet):
    """swap only the three extension modules."""
    model = fx.seeded_model(cuda)
    leg = refnet.Reference(model.state_dict(), cuda, backend="legacy")
    new = refnet.Reference(model.state_dict(), cuda, backend="b200")
    pts = fx.scenes().to(cuda)
    a, b = leg.forward(pts), new.forward(pts)
    # identical Python, identical cuDNN layers; the extension ops are bit-exact twins -> every output is EQUAL
    for k in ("rpn_cls", "rpn_reg", "backbone_features", "seg_result", "rois", "roi_scores_raw", "rcnn_cls", "rcnn_reg"):
        assert torch.equal(a[k], b[k]), k
    da, db = leg.eval_batch(fx.scenes()), new.eval_batch(fx.scenes())
    assert len(da) == len(db)
    for (ba, sa), (bb, sb) in zip(da, db):
        assert np.array_equal(ba, bb) and np.array_equal(sa, sb)


def test_reference_op_modules_on_the_b200_stubs(cuda, refnet):
    """the reference's pointnet2_utils.py / pointnet2_modules.py (op level) on the package's pointnet2_cuda stub against
    the same files on the reference kernels: FPS, ball query, grouping, three_nn / interpolate, one MSG SA module."""
    syn = load("synthetic")
    xyz = torch.from_numpy(syn.make_clouds("lidar", 2, 4096, seed=5)).to(cuda)
    feats = torch.randn((2, 16, 4096), generator=torch.Generator().manual_seed(1)).to(cuda)
    outs = {}
    for backend in ("legacy", "b200"):
        with refnet.reference_imports(backend):
            import pointnet2_lib.pointnet2.pointnet2_utils as pu
            import pointnet2_lib.pointnet2.pointnet2_modules as pm
            torch.manual_seed(0)
            sa = pm.PointnetSAModuleMSG(npoint=512, radii=[0.5, 1.0], nsamples=[16, 32], mlps=[[16, 16, 32], [16, 32, 64]],
                                        use_xyz=True, bn=True).cuda().eval()
            with torch.no_grad():
                idx = pu.furthest_point_sample(xyz, 512)
                new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
                bq = pu.ball_query(1.0, 32, xyz, new_xyz)
                grouped = pu.grouping_operation(feats, bq)
                dist, nn_idx = pu.three_nn(xyz, new_xyz)
                w = 1.0 / (dist + 1e-8)
                w = w / w.sum(dim=2, keepdim=True)
                interp = pu.three_interpolate(feats[:, :, :512].contiguous(), nn_idx, w)
                sa_xyz, sa_feat = sa(xyz, feats)
            outs[backend] = (idx, new_xyz, bq, grouped, dist, nn_idx, interp, sa_xyz, sa_feat)
    for i, (x, y) in enumerate(zip(outs["legacy"], outs["b200"])):
        assert torch.equal(x, y), "output %d differs between the reference kernels and the b200 stubs" % i
