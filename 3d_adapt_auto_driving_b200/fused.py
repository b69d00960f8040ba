"""Thin host-side helpers that enqueue the fused sm_100a kernels on point-major tensors.

Everything here is launch plumbing: argument marshalling for the C-ABI (include/pn2_b200.h),
weight folding/packing caches, output allocation.  No arithmetic is done in torch except the
three_nn weight normalisation, which the reference itself does in torch
(pointnet2_modules.py:140-142).
"""
import torch
import torch.nn as nn

from . import cabi
from .cabi import i32, f32, ptr
from . import pytorch_utils as pt_utils
import ctypes


def _i64(x):
    return ctypes.c_longlong(int(x))


def _rows2d(x):
    """(..., C) tensor whose last dim is contiguous and whose leading dims are jointly
    contiguous rows with one stride -> (data_ptr tensor, rows, ld, C)."""
    c = x.shape[-1]
    if x.stride(-1) != 1 and c > 1:
        raise cabi.Pn2Error("last dimension must be contiguous")
    x2 = x.reshape(-1, c) if x.is_contiguous() else x
    if x2.dim() != 2:
        raise cabi.Pn2Error("expected a 2-D row view; make the tensor contiguous first")
    return x2, x2.shape[0], (x2.stride(0) if x2.shape[0] > 1 else max(c, x2.stride(0))), c


class PackedLayer:
    """Folded (conv + eval BN) layer: w (cout, kpad) zero-padded to a multiple of 4, bias, relu."""

    def __init__(self, w, b, relu):
        cout, cin = w.shape
        kpad = (cin + 3) // 4 * 4
        wp = torch.zeros((cout, kpad), dtype=torch.float32, device=w.device)
        wp[:, :cin] = w
        self.w, self.b, self.relu, self.cin, self.cout, self.ldw = wp, b.contiguous(), bool(relu), cin, cout, kpad

    @staticmethod
    def from_block(block):
        return PackedLayer(*pt_utils.fold_layer(block))


def pack_sequential(seq):
    """Conv blocks of an nn.Sequential (SharedMLP / head); Dropout is the identity at inference."""
    out = []
    for m in seq.children():
        if isinstance(m, nn.Dropout):
            continue
        out.append(PackedLayer.from_block(m))
    return out


def linear(x, layer, out=None, pool=1, res=None, relu=None):
    """y = act(x @ W^T + b [+ res]) on rows; x (rows, cin) view, out optional (rows/pool, cout) view."""
    x2, rows, ldx, cin = _rows2d(x)
    if cin != layer.cin:
        raise cabi.Pn2Error("linear: input has %d channels, layer expects %d" % (cin, layer.cin))
    if rows % pool:
        raise cabi.Pn2Error("linear: rows not divisible by pool")
    if out is None:
        out = torch.empty((rows // pool, layer.cout), dtype=torch.float32, device=x.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows // pool and oc == layer.cout, (orows, rows, pool, oc, layer.cout)
    rp, ldr = ptr(None), 0
    if res is not None:
        r2, rrows, ldr, rc = _rows2d(res)
        assert rrows == rows and rc == layer.cout
        rp = ptr(r2)
    use_relu = layer.relu if relu is None else relu
    cabi.call("pn2_linear_f32", ptr(x2), i32(ldx), ptr(layer.w), i32(layer.ldw), ptr(layer.b), rp, i32(ldr), ptr(o2),
              i32(ldy), _i64(rows), i32(cin), i32(layer.cout), i32(1 if use_relu else 0), i32(pool),
              work=2.0 * rows * cin * layer.cout)
    return out


def sa_group_linear(h, idx, xyz, centres, wxyz, layer, out=None, pool=1):
    """second SA layer with the gather + split first layer fused in (pn2_sa_group_linear_f32)."""
    B, M, ns = idx.shape
    N = xyz.shape[1]
    h2, hrows, ldh, c1 = _rows2d(h)
    assert hrows == B * N and c1 == layer.cin
    rows = B * M * ns
    if out is None:
        out = torch.empty((rows // pool, layer.cout), dtype=torch.float32, device=h.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows // pool and oc == layer.cout
    cabi.call("pn2_sa_group_linear_f32", ptr(h2), i32(ldh), ptr(idx), ptr(xyz), ptr(centres), ptr(wxyz), ptr(layer.w),
              i32(layer.ldw), ptr(layer.b), ptr(o2), i32(ldy), i32(B), i32(N), i32(M), i32(ns), i32(c1),
              i32(layer.cout), i32(1 if layer.relu else 0), i32(pool), work=2.0 * rows * c1 * (layer.cout + 3))
    return out


def three_interpolate_pm(feats_pm, idx, weight, out):
    """feats (B,m,C) point-major, idx/weight (B,n,3) -> writes out rows (B*n, >=C) cols [0,C)."""
    B, m, C = feats_pm.shape
    n = idx.shape[1]
    f2, _, ldf, _ = _rows2d(feats_pm)
    o2, orows, ldo, oc = _rows2d(out)
    assert orows == B * n and oc == C
    cabi.call("pn2_three_interpolate_pm_f32", ptr(f2), i32(ldf), ptr(idx), ptr(weight), ptr(o2), i32(ldo), i32(B),
              i32(C), i32(m), i32(n), work=4.0 * B * n * C * 4 + 24.0 * B * n)
    return out


def ball_query_dual(xyz, new_xyz, r0, ns0, r1, ns1):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    i0 = torch.zeros((B, M, ns0), dtype=torch.int32, device=xyz.device)
    i1 = torch.zeros((B, M, ns1), dtype=torch.int32, device=xyz.device)
    cabi.call("pn2_ball_query_dual_f32", ptr(new_xyz), ptr(xyz), ptr(i0), ptr(i1), i32(B), i32(N), i32(M), f32(r0),
              i32(ns0), f32(r1), i32(ns1), work=12.0 * B * M * N)
    return i0, i1


def fps_gather(xyz, npoint):
    """FPS indices (B,npoint) int32 and the sampled centres (B,npoint,3)."""
    B, N, _ = xyz.shape
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    cabi.call("pn2_fps_f32", ptr(xyz), ptr(None), ptr(idx), i32(B), i32(N), i32(npoint),
              work=16.0 * B * max(npoint - 1, 0) * N)
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    return idx, new_xyz
