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
