"""GPU parity of the iou3d (rotated BEV overlap / IoU / NMS) and roipool3d kernels, called through
the reference-shaped modules (iou3d_utils / iou3d_cuda / roipool3d_utils -> C-ABI), against
  (1) golden vectors produced on a B200 by the reference's own kernels (tests/golden/*.npz,
      tools/make_goldens.py), and
  (2) those kernels themselves (oracle/_ref/libpn2_legacy.so) at the proposal layer's sizes.
Overlap areas, IoUs, keep lists and pooled indices must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import load

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
synthetic = load("synthetic")


def _bev_boxes(seed, n, spread=20.0):
    rng = np.random.RandomState(seed)
    cx = rng.uniform(-spread, spread, n); cz = rng.uniform(0, 2 * spread, n)
    l = rng.uniform(3.0, 4.8, n); w = rng.uniform(1.4, 2.0, n)
    ry = rng.uniform(-np.pi, np.pi, n)
    dup = rng.randint(0, n, n // 2)
    cx[: n // 2] = cx[dup] + rng.normal(0, 0.15, n // 2)
    cz[: n // 2] = cz[dup] + rng.normal(0, 0.15, n // 2)
    ry[: n // 2] = ry[dup] + rng.normal(0, 0.05, n // 2)
    return np.stack([cx - l / 2, cz - w / 2, cx + l / 2, cz + w / 2, ry], 1).astype(np.float32)


def test_overlap_and_iou_match_reference_goldens(cuda):
    g = np.load(os.path.join(GOLD, "iou3d_legacy.npz"))
    ic = load("iou3d_cuda")
    a, b = torch.from_numpy(g["a"]).to(cuda), torch.from_numpy(g["b"]).to(cuda)
    ov = torch.zeros((a.shape[0], b.shape[0]), device=cuda)
    iou = torch.zeros_like(ov)
    ic.boxes_overlap_bev_gpu(a, b, ov)
    ic.boxes_iou_bev_gpu(a, b, iou)
    assert np.array_equal(ov.cpu().numpy(), g["overlap"])
    assert np.array_equal(iou.cpu().numpy(), g["iou"])
    assert torch.equal(load("iou3d_utils").boxes_iou_bev(a, b), iou)


def test_nms_keep_lists_match_reference_goldens(cuda):
    g = np.load(os.path.join(GOLD, "iou3d_legacy.npz"))
    ic = load("iou3d_cuda")
    boxes = torch.from_numpy(g["nms_boxes"]).to(cuda)
    for thr in (0.1, 0.8):
        for name, fn in (("rot", ic.nms_gpu), ("nrm", ic.nms_normal_gpu)):
            keep = torch.zeros(boxes.shape[0], dtype=torch.int64)  # CPU LongTensor, the reference contract
            n = fn(boxes, keep, thr)
            assert np.array_equal(keep[:n].numpy(), g["keep_%s_%g" % (name, thr)]), (name, thr)


@pytest.mark.parametrize("normal", [True, False])
def test_nms_proposal_sizes_vs_legacy_and_early_stop(cuda, legacy, normal):
    # proposal_layer.py:58-119: 6300 / 2700 pre-NMS boxes, thresh 0.8, first 70 / 30 survivors
    ic = load("iou3d_cuda")
    for n, post in ((6300, 70), (2700, 30), (65, 70)):
        boxes = torch.from_numpy(_bev_boxes(n, n, spread=35.0)).to(cuda)
        ref = legacy.greedy_from_mask(legacy.nms_mask(boxes, 0.8, normal=normal).cpu(), n)
        keep, num = ic.nms_device(boxes, 0.8, rotated=not normal)
        assert int(num) == len(ref) and np.array_equal(keep[0, :len(ref)].cpu().numpy(), ref)
        keep, num = ic.nms_device(boxes, 0.8, rotated=not normal, max_keep=post)
        k = min(post, len(ref))
        assert int(num) == k and np.array_equal(keep[0, :k].cpu().numpy(), ref[:k])


def test_nms_batched_with_device_counts(cuda, legacy):
    ic = load("iou3d_cuda")
    P, stride = 5, 1000
    boxes = torch.stack([torch.from_numpy(_bev_boxes(10 + p, stride)) for p in range(P)]).to(cuda)
    counts = torch.tensor([1000, 0, 1, 333, 64], dtype=torch.int32, device=cuda)
    keep, num = ic.nms_device(boxes, 0.1, rotated=True, max_keep=stride, counts=counts)
    for p in range(P):
        n = int(counts[p])
        ref = legacy.greedy_from_mask(legacy.nms_mask(boxes[p, :max(n, 1)].contiguous(), 0.1).cpu(), n) if n else np.zeros(0, np.int64)
        assert int(num[p]) == len(ref), p
        assert np.array_equal(keep[p, :len(ref)].cpu().numpy(), ref), p


def test_nms_rotated_dense_mode_and_large_problems(cuda, legacy):
    """rotated problems of <= 128 boxes take the all-pairs bit-matrix path, larger ones the lazy rows;
    problems beyond the shared-memory staging budget (9830 boxes) read the boxes from global memory"""
    ic = load("iou3d_cuda")
    P, stride = 9, 129
    sizes = [0, 1, 2, 37, 64, 100, 127, 128, 129]
    boxes = torch.stack([torch.from_numpy(_bev_boxes(70 + p, stride, spread=6.0)) for p in range(P)]).to(cuda)
    counts = torch.tensor(sizes, dtype=torch.int32, device=cuda)
    for thr in (0.1, 0.7):
        for max_keep in (stride, 10):
            keep, num = ic.nms_device(boxes, thr, rotated=True, max_keep=max_keep, counts=counts)
            for p, n in enumerate(sizes):
                ref = legacy.greedy_from_mask(legacy.nms_mask(boxes[p, :max(n, 1)].contiguous(), thr).cpu(), n) if n else np.zeros(0, np.int64)
                k = min(max_keep, len(ref))
                assert int(num[p]) == k, (thr, max_keep, n)
                assert np.array_equal(keep[p, :k].cpu().numpy(), ref[:k]), (thr, max_keep, n)
    big = torch.from_numpy(_bev_boxes(5, 12000, spread=60.0)).to(cuda)
    ref = legacy.greedy_from_mask(legacy.nms_mask(big, 0.8, normal=True).cpu(), 12000)
    keep, num = ic.nms_device(big, 0.8, rotated=False, max_keep=200)
    assert int(num) == 200 and np.array_equal(keep[0, :200].cpu().numpy(), ref[:200])


def test_nms_utils_returns_original_indices(cuda, legacy):
    iu = load("iou3d_utils")
    boxes = torch.from_numpy(_bev_boxes(7, 500)).to(cuda)
    scores = torch.rand(500, device=cuda)
    order = scores.sort(0, descending=True)[1]
    ref = legacy.greedy_from_mask(legacy.nms_mask(boxes[order].contiguous(), 0.3).cpu(), 500)
    got = iu.nms_gpu(boxes, scores, 0.3)
    assert got.is_cuda and got.dtype == torch.int64
    assert torch.equal(got.cpu(), order.cpu()[torch.from_numpy(ref)])
    assert iu.nms_gpu(boxes[:0], scores[:0], 0.3).numel() == 0


def test_boxes_iou3d_vs_legacy_overlap(cuda, legacy):
    iu, ku = load("iou3d_utils"), load("kitti_utils")
    rng = np.random.RandomState(0)
    def boxes3d(n):
        return torch.from_numpy(np.stack([rng.uniform(-10, 10, n), rng.uniform(0.5, 2, n), rng.uniform(5, 40, n),
                                          rng.uniform(1.3, 1.8, n), rng.uniform(1.4, 1.9, n), rng.uniform(3, 5, n),
                                          rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)).to(cuda)
    a, b = boxes3d(90), boxes3d(40)
    b[:20] = a[:20] + 0.05
    got = iu.boxes_iou3d_gpu(a, b)
    ov = legacy.boxes_overlap_bev(ku.boxes3d_to_bev_torch(a).contiguous(), ku.boxes3d_to_bev_torch(b).contiguous())
    a_min, a_max = (a[:, 1] - a[:, 3]).view(-1, 1), a[:, 1].view(-1, 1)
    b_min, b_max = (b[:, 1] - b[:, 3]).view(1, -1), b[:, 1].view(1, -1)
    oh = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    o3 = ov * oh
    va = (a[:, 3] * a[:, 4] * a[:, 5]).view(-1, 1); vb = (b[:, 3] * b[:, 4] * b[:, 5]).view(1, -1)
    assert torch.equal(got, o3 / torch.clamp(va + vb - o3, min=1e-7))      # the one-launch kernel == reference composition
    assert torch.equal(got, iu.boxes_iou3d_torch(a, b))
    assert got.max() > 0.5
    assert iu.boxes_iou3d_gpu(a[:0], b).shape == (0, 40) and iu.boxes_iou3d_gpu(a, b[:0]).shape == (90, 0)
    assert torch.equal(iu.boxes_iou3d_gpu(a[3:4], a[3:4]), iu.boxes_iou3d_torch(a[3:4], a[3:4]))


def test_roipool3d_matches_reference_golden(cuda):
    import hashlib
    g = np.load(os.path.join(GOLD, "roipool3d_legacy.npz"))
    xyz = torch.from_numpy(synthetic.make_clouds("lidar", 2, 16384, seed=1024)).to(cuda)
    feat = torch.from_numpy(np.random.RandomState(3).randn(2, 16384, 5).astype(np.float32)).to(cuda)
    boxes = torch.from_numpy(g["boxes"]).to(cuda)
    pooled = torch.zeros((2, boxes.shape[1], 512, 8), device=cuda)
    empty = torch.zeros((2, boxes.shape[1]), dtype=torch.int32, device=cuda)
    load("roipool3d_cuda").forward(xyz, boxes, feat, pooled, empty)
    assert np.array_equal(empty.cpu().numpy(), g["empty"])
    assert np.array_equal(pooled[..., :3].cpu().numpy(), g["pooled_xyz"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(pooled.cpu().numpy().tobytes()).digest(), np.uint8), g["sha"])


def test_roipool3d_utils_full_size_vs_legacy(cuda, legacy):
    # RCNN input of config 4: 16384 points x 130 features, 100 ROIs per scene (some all-zero)
    ku = load("kitti_utils")
    B, N, M, C = 3, 16384, 100, 130
    xyz_h = synthetic.make_clouds("lidar", B, N, seed=666)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    feat = torch.randn((B, N, C), device=cuda)
    rng = np.random.RandomState(1)
    rois = np.zeros((B, M, 7), np.float32)
    for b in range(B):
        for m in range(80):
            p = xyz_h[b, rng.randint(0, N)]
            rois[b, m] = [p[0], p[1] + 0.8, p[2], 1.5, 1.6, 3.9, rng.uniform(-np.pi, np.pi)]
    rois_t = torch.from_numpy(rois).to(cuda)
    pooled, empty = load("roipool3d_utils").roipool3d_gpu(xyz, feat, rois_t, 1.0, sampled_pt_num=512)
    enlarged = ku.enlarge_box3d(rois_t.view(-1, 7), 1.0).view(B, M, 7).contiguous()
    ref_pooled, ref_empty = legacy.roipool3d(xyz, feat, enlarged, sampled=512)
    assert torch.equal(empty, ref_empty)
    assert torch.equal(pooled, ref_pooled)
    assert pooled.shape == (B, M, 512, 3 + C) and int(empty.sum()) < B * M


def test_roipool3d_split_padded_rows_equal_plain_pooling(cuda):
    """pn2_roipool3d_split_f32 (two feature sources, padded rows) pools the same points in the same order"""
    cabi = load("cabi")
    B, N, M, S = 2, 16384, 40, 512
    xyz_h = synthetic.make_clouds("lidar", B, N, seed=31)
    xyz = torch.from_numpy(xyz_h).to(cuda)
    head = torch.randn((B, N, 2), device=cuda)
    wide = torch.randn((B, N, 128), device=cuda)
    rng = np.random.RandomState(2)
    rois = np.zeros((B, M, 7), np.float32)
    for b in range(B):
        for m in range(M - 5):            # the last five boxes stay all-zero (empty)
            p = xyz_h[b, rng.randint(0, N)]
            rois[b, m] = [p[0], p[1] + 0.8, p[2], 2.5, 2.6, 4.9, rng.uniform(-np.pi, np.pi)]
    boxes = torch.from_numpy(rois).to(cuda)
    ref = torch.zeros((B, M, S, 3 + 130), device=cuda)
    ref_empty = torch.zeros((B, M), dtype=torch.int32, device=cuda)
    load("roipool3d_cuda").forward(xyz, boxes, torch.cat((head, wide), dim=2).contiguous(), ref, ref_empty)
    got = torch.zeros((B, M, S, 136), device=cuda)
    empty = torch.zeros((B, M), dtype=torch.int32, device=cuda)
    cabi.call("pn2_roipool3d_split_f32", cabi.ptr(xyz), cabi.ptr(boxes), cabi.ptr(head), cabi.i32(2), cabi.ptr(wide),
              cabi.i32(128), cabi.ptr(got), cabi.i32(136), cabi.i32(8), cabi.ptr(empty), cabi.i32(B), cabi.i32(N),
              cabi.i32(M), cabi.i32(S))
    assert torch.equal(empty, ref_empty) and int(empty.sum()) >= 5 * B
    assert torch.equal(got[..., :5], ref[..., :5]) and torch.equal(got[..., 8:], ref[..., 5:])
    assert float(got[..., 5:8].abs().max()) == 0.0
