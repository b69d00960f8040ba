"""numpy-facing wrapper of oracle/libpn2_oracle.so (the C restatement in pn2_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  Inputs and outputs are
numpy arrays on the host; every function cites the reference kernel it restates in the C file.
"""
import ctypes
import os

import numpy as np

from . import build_ref

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_ref.build_oracle())
        _lib.orc_opt_n_threads.restype = ctypes.c_int
        _lib.orc_nms_normal.restype = ctypes.c_int
        _lib.orc_nms_rotated.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n):
    return lib().orc_opt_n_threads(int(n))


def fps(xyz, npoint, temp=None):
    """xyz (B,N,3) -> idx (B,npoint) int32, temp (B,N) after the call."""
    xyz, px = _f(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, np.float32) if temp is None else np.array(temp, np.float32, copy=True)
    idx = np.zeros((B, npoint), np.int32)
    lib().orc_fps(px, temp.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), B, N, int(npoint))
    return idx, temp


def fps_pruned_model(xyz, npoint, order, cell=128, temp=None):
    """CPU model of the algorithm of csrc/fps_cells.cu (oracle/fps_pruned_model.c) for ONE cloud: xyz (N,3), order (N) a
    permutation -> idx (npoint) int32, temp (N) after the call, number of touched (round, cell) pairs."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    order = np.ascontiguousarray(order, np.int32)
    N = xyz.shape[0]
    temp = np.full((N,), 1e10, np.float32) if temp is None else np.array(temp, np.float32, copy=True)
    idx = np.zeros((npoint,), np.int32)
    fn = lib().orc_fps_pruned_model
    fn.restype = ctypes.c_longlong
    touched = fn(xyz.ctypes.data_as(ctypes.c_void_p), order.ctypes.data_as(ctypes.c_void_p), N, int(npoint), int(cell),
                 temp.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p))
    return idx, temp, int(touched)


def gather_points(points, idx):
    points, pp = _f(points); idx, pi = _i(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), np.float32)
    lib().orc_gather_points(pp, pi, out.ctypes.data_as(ctypes.c_void_p), B, C, N, M)
    return out


def gather_points_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out); idx, pi = _i(idx)
    B, C, M = grad_out.shape
    gp = np.zeros((B, C, N), np.float32)
    lib().orc_gather_points_grad(pg, pi, gp.ctypes.data_as(ctypes.c_void_p), B, C, N, M)
    return gp


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz); new_xyz, pn = _f(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(pn, px, idx.ctypes.data_as(ctypes.c_void_p), B, N, M, ctypes.c_float(radius), int(nsample))
    return idx


def group_points(points, idx):
    points, pp = _f(points); idx, pi = _i(idx)
    B, C, N = points.shape
    _, M, ns = idx.shape
    out = np.empty((B, C, M, ns), np.float32)
    lib().orc_group_points(pp, pi, out.ctypes.data_as(ctypes.c_void_p), B, C, N, M, ns)
    return out


def group_points_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out); idx, pi = _i(idx)
    B, C, M, ns = grad_out.shape
    gp = np.zeros((B, C, N), np.float32)
    lib().orc_group_points_grad(pg, pi, gp.ctypes.data_as(ctypes.c_void_p), B, C, N, M, ns)
    return gp


def three_nn(unknown, known):
    """-> (dist2 (B,n,3) SQUARED, idx (B,n,3))"""
    unknown, pu = _f(unknown); known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((B, n, 3), np.float32)
    idx = np.empty((B, n, 3), np.int32)
    lib().orc_three_nn(pu, pk, d2.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), B, n, m)
    return d2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points); idx, pi = _i(idx); weight, pw = _f(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), np.float32)
    lib().orc_three_interpolate(pp, pi, pw, out.ctypes.data_as(ctypes.c_void_p), B, C, m, n)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out); idx, pi = _i(idx); weight, pw = _f(weight)
    B, C, n = grad_out.shape
    gp = np.zeros((B, C, m), np.float32)
    lib().orc_three_interpolate_grad(pg, pi, pw, gp.ctypes.data_as(ctypes.c_void_p), B, C, n, m)
    return gp


def nms_normal(boxes_sorted, thresh):
    """boxes (n,5) already sorted by score -> keep indices (into the sorted list)."""
    boxes, pb = _f(boxes_sorted)
    n = boxes.shape[0]
    keep = np.zeros((max(n, 1),), np.int64)
    k = lib().orc_nms_normal(pb, keep.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_float(thresh))
    return keep[:k].copy()


def roipool3d(xyz, feat, boxes3d_enlarged, sampled=512, trig=None):
    xyz, px = _f(xyz); feat, pf = _f(feat); boxes, pb = _f(boxes3d_enlarged)
    B, N, _ = xyz.shape
    M = boxes.shape[1]
    C = feat.shape[2]
    pooled = np.zeros((B, M, sampled, 3 + C), np.float32)
    empty = np.zeros((B, M), np.int32)
    pt = ctypes.c_void_p(0)
    if trig is not None:
        trig, pt = _f(trig)
    lib().orc_roipool3d(px, pb, pf, pooled.ctypes.data_as(ctypes.c_void_p), empty.ctypes.data_as(ctypes.c_void_p), pt,
                        B, N, M, C, int(sampled))
    return pooled, empty


# ---- geom_oracle.c: rotated BEV overlap / IoU / NMS (iou3d_kernel.cu + iou3d.cpp) ----
def boxes_overlap_bev(a, b, iou=False):
    a, pa = _f(a); b, pb = _f(b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_overlap_bev(pa, a.shape[0], pb, b.shape[0], out.ctypes.data_as(ctypes.c_void_p), 1 if iou else 0)
    return out


def nms_rotated(boxes_sorted, thresh):
    boxes, pb = _f(boxes_sorted)
    n = boxes.shape[0]
    keep = np.zeros((max(n, 1),), np.int64)
    k = lib().orc_nms_rotated(pb, keep.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_float(thresh))
    return keep[:k].copy()


def rotate_iou_eval(boxes, query_boxes, criterion=-1, return_npts=False):
    """geom_oracle.c restatement of evaluate/rotate_iou.py's kernel -> (N,K) float32 [, vertex counts]."""
    b, pb = _f(boxes); q, pq = _f(query_boxes)
    n, k = b.shape[0], q.shape[0]
    out = np.zeros((n, k), np.float32)
    npts = np.zeros((n, k), np.int32)
    lib().orc_rotate_iou_eval(pb, n, pq, k, out.ctypes.data_as(ctypes.c_void_p), int(criterion),
                              npts.ctypes.data_as(ctypes.c_void_p))
    return (out, npts) if return_npts else out
