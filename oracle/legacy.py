"""torch-facing door onto oracle/_ref/libpn2_legacy.so = the reference's OWN CUDA kernels
compiled unchanged (oracle/build_ref.py).  TEST / BASELINE INFRASTRUCTURE ONLY.

Needs a GPU.  Used by the -m gpu tests to pin both the CPU oracle and the new kernels to the
real reference, by tools/make_goldens.py to produce tests/golden/*.npz, and by bench.py to
time the legacy-CUDA path next to the new one.
"""
import ctypes
import os

import torch

from . import build_ref

_lib = None


def available():
    return build_ref.build_legacy() is not None


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build_legacy()
        if path is None:
            raise RuntimeError("oracle/_ref/libpn2_legacy.so is absent and /root/reference is not mounted")
        _lib = ctypes.CDLL(path)
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def fps(xyz, npoint, temp=None):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, device=xyz.device) if temp is None else temp
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    lib().legacy_fps(B, N, int(npoint), _p(xyz), _p(temp), _p(idx), _s())
    return idx, temp


def gather(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), device=points.device)
    lib().legacy_gather(B, C, N, M, _p(points), _p(idx), _p(out), _s())
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)
    lib().legacy_ball_query(B, N, M, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def group(points, idx):
    B, C, N = points.shape
    _, M, ns = idx.shape
    out = torch.empty((B, C, M, ns), device=points.device)
    lib().legacy_group(B, C, N, M, ns, _p(points), _p(idx), _p(out), _s())
    return out


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty((B, n, 3), device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    lib().legacy_three_nn(B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_interpolate(points, idx, weight):
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), device=points.device)
    lib().legacy_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())
    return out


# iou3d / roipool3d launchers run on the legacy default stream: synchronise around them.
def boxes_overlap_bev(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), device=a.device)
    torch.cuda.synchronize()
    lib().legacy_boxes_overlap_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    torch.cuda.synchronize()
    return out


def boxes_iou_bev(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), device=a.device)
    torch.cuda.synchronize()
    lib().legacy_boxes_iou_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    torch.cuda.synchronize()
    return out


def nms_mask(boxes, thresh, normal=False):
    n = boxes.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=boxes.device)
    torch.cuda.synchronize()
    fn = lib().legacy_nms_normal_mask if normal else lib().legacy_nms_mask
    fn(_p(boxes), _p(mask), n, ctypes.c_float(thresh))
    torch.cuda.synchronize()
    return mask


def greedy_from_mask(mask_cpu, n):
    """host greedy pass of iou3d.cpp:100-116 over the u64 suppression masks."""
    import numpy as np
    m = mask_cpu.numpy().view(np.uint64)
    cb = m.shape[1]
    remv = np.zeros((cb,), np.uint64)
    keep = []
    for i in range(n):
        nb, ib = divmod(i, 64)
        if not (int(remv[nb]) >> ib) & 1:
            keep.append(i)
            remv[nb:] |= m[i, nb:]
    return np.asarray(keep, np.int64)


def roipool3d(xyz, feat, boxes_enlarged, sampled=512):
    B, N, _ = xyz.shape
    M = boxes_enlarged.shape[1]
    C = feat.shape[2]
    pooled = torch.zeros((B, M, sampled, 3 + C), device=xyz.device)
    empty = torch.zeros((B, M), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    lib().legacy_roipool3d(B, N, M, C, int(sampled), _p(xyz), _p(boxes_enlarged), _p(feat), _p(pooled), _p(empty))
    torch.cuda.synchronize()
    return pooled, empty


# ------------------------------------------------------------------------------------------
# The legacy-CUDA arm of bench.py (--impl reference): route the product's reference-shaped
# Python modules (fused = False: the reference's own composition) onto the REFERENCE kernels.
# Runs in its own process; nothing is restored.
# ------------------------------------------------------------------------------------------
def nms_reference(boxes, scores, thresh, normal, max_keep=None):
    """iou3d_utils.py:56-87 as the reference executes it: sort on the device, mask kernel,
    blocking D2H of the (n, n/64) u64 matrix, greedy pass on the host (C, like iou3d.cpp),
    keep list re-uploaded."""
    import numpy as np
    from . import oracle as orc
    order = scores.sort(0, descending=True)[1]
    b = boxes[order].contiguous()
    n = b.shape[0]
    if n == 0:
        return order[:0]
    mask = nms_mask(b, thresh, normal=normal).cpu().numpy()
    keep = np.zeros((n,), np.int64)
    orc.lib().orc_nms_greedy_from_mask.restype = ctypes.c_int
    k = orc.lib().orc_nms_greedy_from_mask(mask.ctypes.data_as(ctypes.c_void_p), n, keep.ctypes.data_as(ctypes.c_void_p))
    return order[torch.from_numpy(keep[:k]).to(boxes.device)].contiguous()


def install(pkg_name):
    """Monkey-patch <pkg>.pointnet2_cuda / iou3d_utils / roipool3d_cuda to call libpn2_legacy.so."""
    import importlib
    L = lib()
    p2 = importlib.import_module(pkg_name + ".pointnet2_cuda")
    iu = importlib.import_module(pkg_name + ".iou3d_utils")
    rp = importlib.import_module(pkg_name + ".roipool3d_cuda")

    def fps_w(b, n, m, points, temp, idx):
        L.legacy_fps(b, n, m, _p(points), _p(temp), _p(idx), _s()); return 1

    def gather_w(b, c, n, npoints, points, idx, out):
        L.legacy_gather(b, c, n, npoints, _p(points), _p(idx), _p(out), _s()); return 1

    def bq_w(b, n, m, radius, nsample, new_xyz, xyz, idx):
        L.legacy_ball_query(b, n, m, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s()); return 1

    def group_w(b, c, n, npoints, nsample, points, idx, out):
        L.legacy_group(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), _s()); return 1

    def nn_w(b, n, m, unknown, known, dist2, idx):
        L.legacy_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), _s())

    def interp_w(b, c, m, n, points, idx, weight, out):
        L.legacy_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())

    p2.furthest_point_sampling_wrapper = fps_w
    p2.gather_points_wrapper = gather_w
    p2.ball_query_wrapper = bq_w
    p2.group_points_wrapper = group_w
    p2.three_nn_wrapper = nn_w
    p2.three_interpolate_wrapper = interp_w

    iu.nms_gpu = lambda boxes, scores, thresh, max_keep=None: nms_reference(boxes, scores, thresh, False)
    iu.nms_normal_gpu = lambda boxes, scores, thresh, max_keep=None: nms_reference(boxes, scores, thresh, True)

    def roipool_fw(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
        B, N, _ = xyz.shape
        # the reference launcher runs on the legacy default stream and cudaMallocs inside
        torch.cuda.current_stream().synchronize()
        L.legacy_roipool3d(B, N, boxes3d.size(1), pts_feature.size(2), pooled_features.size(2), _p(xyz), _p(boxes3d),
                           _p(pts_feature), _p(pooled_features), _p(pooled_empty_flag))
        return 1

    rp.forward = roipool_fw
