"""Thin host-side helpers that enqueue the fused sm_100a kernels on point-major tensors.

Everything here is launch plumbing: argument marshalling for the C-ABI (include/pn2_b200.h),
weight folding/packing caches, output allocation.  No arithmetic is done in torch except the
three_nn weight normalisation, which the reference itself does in torch
(pointnet2_modules.py:140-142).
"""
import os

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


# Which engine runs the shared-MLP layers: "tc" = tcgen05 tensor cores with BF16x3 split operands
# (default; ~1e-5 of the tensor scale per layer), "ffma" = exact-fp32 CUDA-core kernels.
MLP_ENGINE = os.environ.get("PN2_MLP", "tc")
_TC_POOLS = (1, 16, 32, 64, 128)


def set_mlp_engine(name):
    global MLP_ENGINE
    if name not in ("tc", "ffma"):
        raise ValueError(name)
    MLP_ENGINE = name


class PackedLayer:
    """Folded (conv + eval BN) layer: w (cout, kpad) zero-padded to a multiple of 4, bias, relu;
    `.tc` is the same layer packed for the tensor-core kernel (built on first use)."""

    def __init__(self, w, b, relu):
        cout, cin = w.shape
        kpad = (cin + 3) // 4 * 4
        wp = torch.zeros((cout, kpad), dtype=torch.float32, device=w.device)
        wp[:, :cin] = w
        self.w, self.b, self.relu, self.cin, self.cout, self.ldw = wp, b.contiguous(), bool(relu), cin, cout, kpad
        self._tc = None
        self._w3t = None

    @property
    def w3t(self):
        """(hi, lo) int32 tensors (cout, cin / 2): the layer as the tensor-memory A operand of the transposed
        last SA layer (csrc/sa_fused_t_tc.cu): bf16 pairs (k even in the low half), row = output channel."""
        if self._w3t is None:
            self._w3t = pack_w3t(self.w[:, :self.cin])
        return self._w3t

    @property
    def tc(self):
        if self._tc is None:
            self._tc = PackedLayerTC(self.w[:, :self.cin], self.b, self.relu)
        return self._tc

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
    use_relu = layer.relu if relu is None else relu
    if MLP_ENGINE == "tc" and pool in _TC_POOLS and (pool == 1 or use_relu):
        return linear_tc(x, layer.tc, out=out, pool=pool, res=res, relu=use_relu)
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
    if MLP_ENGINE == "tc" and pool in _TC_POOLS and (pool == 1 or layer.relu) and layer.cin <= 512:
        return sa_group_linear_tc(h, idx, xyz, centres, wxyz, layer.tc, out=out, pool=pool)
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


# association torch.sum(dist_recip, dim=2) uses for its three terms (its reduction kernel gives the even and the odd elements
# to two threads): 1 = (r0 + r2) + r1; pinned bit for bit by tests/test_pn2_ops_gpu.py
INTERP_SUM_ORDER = 1


def three_interpolate_pm_d2(feats_pm, idx, dist2, out, sum_order=None):
    """three_interpolate_pm with the weights formed in the kernel from three_nn's squared distances (the sqrt, reciprocal,
    sum and division the FP module otherwise runs as five torch launches; bit-identical weights)."""
    B, m, C = feats_pm.shape
    n = idx.shape[1]
    f2, _, ldf, _ = _rows2d(feats_pm)
    o2, orows, ldo, oc = _rows2d(out)
    assert orows == B * n and oc == C
    cabi.call("pn2_three_interpolate_pm_d2_f32", ptr(f2), i32(ldf), ptr(idx), ptr(dist2), ptr(o2), i32(ldo), i32(B),
              i32(C), i32(m), i32(n), i32(INTERP_SUM_ORDER if sum_order is None else sum_order),
              work=4.0 * B * n * C * 4 + 24.0 * B * n)
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
    i0 = torch.empty((B, M, ns0), dtype=torch.int32, device=xyz.device)   # no zero fill: the _fill entry point writes the
    i1 = torch.empty((B, M, ns1), dtype=torch.int32, device=xyz.device)   # zeros of centres without neighbours itself
    order = torch.empty((B, M), dtype=torch.int32, device=xyz.device)    # scratch: Hilbert order of the centres
    hits = torch.empty((2, B, M), dtype=torch.int32, device=xyz.device)  # neighbours found per centre (group_compact)
    cabi.call("pn2_ball_query_culled_fill_f32", ptr(new_xyz), ptr(xyz), ptr(i0), ptr(i1), ptr(order), i32(B), i32(N), i32(M),
              f32(r0), i32(ns0), f32(r1), i32(ns1), ptr(hits[0]), ptr(hits[1]), work=12.0 * B * M * N)
    i0._pn2_hits, i1._pn2_hits = hits[0], hits[1]
    return i0, i1


def ball_query_single(xyz, new_xyz, radius, nsample):
    """pointnet2_utils.ball_query without the zero-fill launch (same lists)."""
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=xyz.device)
    order = torch.empty((B, M), dtype=torch.int32, device=xyz.device)
    hits = torch.empty((B, M), dtype=torch.int32, device=xyz.device)
    cabi.call("pn2_ball_query_culled_fill_f32", ptr(new_xyz), ptr(xyz), ptr(idx), ptr(None), ptr(order), i32(B), i32(N), i32(M),
              f32(radius), i32(nsample), f32(0.0), i32(0), ptr(hits), ptr(None), work=12.0 * B * M * N)
    idx._pn2_hits = hits      # rides along with the lists: group_compact() then skips its counting pass over them
    return idx


# SA levels that sample an already FPS-ordered cloud (every backbone level after the first) first run the exact
# parallel prefix test (csrc/fps.cu: pn2_fps_prefix_check_f32); clouds that pass get arange(npoint) without the round loop.
FPS_PREFIX_CHECK = os.environ.get("PN2_FPS_PREFIX_CHECK", "1") != "0"


# PN2_FPS_CLUSTER = 1 / 2 / 4 / 8 forces the cluster kernel of csrc/fps.cu (every point updated every round) for clouds above
# 4096 points; 0 (default) lets pn2_fps_f32 choose: the pruned one-CTA kernel of csrc/fps_cells.cu up to 16384 points
# (16 x (16384 -> 4096): 1.6 ms on 16 SMs against 2.9 ms on 64 / 3.4 ms on 32, tools/bench_fps_cluster.py).
FPS_CLUSTER = int(os.environ.get("PN2_FPS_CLUSTER", "0"))


def fps_gather(xyz, npoint, fps_ordered=False):
    """FPS indices (B,npoint) int32 and the sampled centres (B,npoint,3).  fps_ordered: the caller knows that xyz is
    the output of a previous furthest point sampling (a hint, never trusted: the prefix test decides per cloud)."""
    B, N, _ = xyz.shape
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    # the sampling kernels write the coordinates of their picks themselves (three stores a round): no cast / gather / copy
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device)
    if fps_ordered and FPS_PREFIX_CHECK and 1 < npoint <= min(N, 4096):
        viol = torch.empty((B,), dtype=torch.int32, device=xyz.device)       # zeroed by the check itself (a memset node)
        dmin = torch.empty((B, npoint), dtype=torch.float32, device=xyz.device)
        cabi.call("pn2_fps_prefix_check_f32", ptr(xyz), ptr(dmin), ptr(viol), i32(B), i32(N), i32(npoint),
                  work=12.0 * B * npoint * N)
        cabi.call("pn2_fps_guarded_xyz_f32", ptr(xyz), ptr(idx), ptr(new_xyz), ptr(viol), i32(B), i32(N), i32(npoint), work=0.0)
    elif FPS_CLUSTER and N > 4096:
        cabi.call("pn2_fps_cluster_f32", ptr(xyz), ptr(None), ptr(idx), i32(B), i32(N), i32(npoint), i32(FPS_CLUSTER),
                  work=16.0 * B * max(npoint - 1, 0) * N)
        new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    else:
        cabi.call("pn2_fps_xyz_f32", ptr(xyz), ptr(idx), ptr(new_xyz), i32(B), i32(N), i32(npoint),
                  work=16.0 * B * max(npoint - 1, 0) * N)
    return idx, new_xyz


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05) path: weights split into bf16 hi/lo and pre-arranged in the exact shared
# memory image the MMA reads (linear_tc.cu): [nchunks][nkb][hi tile | lo tile], each tile
# ntile rows x 64 bf16 (128 B), 16-byte chunk c of row r stored at chunk position c ^ (r & 7).
# ---------------------------------------------------------------------------------------------
def tc_tiling(cout):
    """-> (ntile, nchunks): UMMA N (multiple of 16, <= 256) and the number of N chunks."""
    if cout <= 256:
        return (cout + 15) // 16 * 16, 1
    nchunks = (cout + 255) // 256
    per = (cout + nchunks - 1) // nchunks
    return (per + 15) // 16 * 16, nchunks


def pack_tc(w):
    """w (cout, cin) f32 -> (blob uint8 tensor, ntile, nchunks, nkb)."""
    cout, cin = w.shape
    ntile, nchunks = tc_tiling(cout)
    nkb = (cin + 63) // 64
    wp = torch.zeros((nchunks * ntile, nkb * 64), dtype=torch.float32, device=w.device)
    wp[:cout, :cin] = w
    hi = wp.to(torch.bfloat16)
    lo = (wp - hi.float()).to(torch.bfloat16)
    both = torch.stack((hi, lo), dim=0)                                      # (2, N, K)
    t = both.view(2, nchunks, ntile // 8, 8, nkb, 8, 8)                       # (h, chunk, grp, r, kb, c, e)
    r = torch.arange(8, device=w.device).view(8, 1)
    c = torch.arange(8, device=w.device).view(1, 8)
    src_c = (c ^ r)                                                          # position p holds chunk p ^ r
    t = torch.gather(t, 5, src_c.view(1, 1, 1, 8, 1, 8, 1).expand(2, nchunks, ntile // 8, 8, nkb, 8, 8))
    blob = t.permute(1, 4, 0, 2, 3, 5, 6).contiguous()                       # (chunk, kb, h, grp, r, pos, e)
    return blob.view(torch.uint8).reshape(-1), ntile, nchunks, nkb


def pack_w3t(w):
    """w (cout, cin) f32, cin even -> (hi, lo) int32 (cout padded to a multiple of 128 with zero rows, cin / 2):
    bf16 split x = hi + lo, two k per word."""
    cout, cin = w.shape
    pad = (cout + 127) // 128 * 128
    if pad != cout:
        w = torch.cat((w, torch.zeros((pad - cout, cin), dtype=w.dtype, device=w.device)), dim=0)
    w = w.contiguous()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi.contiguous().view(torch.int32), lo.contiguous().view(torch.int32)


class PackedLayerTC:
    def __init__(self, w, b, relu):
        self.cout, self.cin = w.shape
        self.blob, self.ntile, self.nchunks, self.nkb = pack_tc(w)
        self.b, self.relu = b.contiguous(), bool(relu)


def linear_tc(x, layer, out=None, pool=1, res=None, relu=None):
    """tensor-core twin of linear(): y = act(x @ W^T + b [+ res]) [max over pool rows]."""
    x2, rows, ldx, cin = _rows2d(x)
    if cin != layer.cin:
        raise cabi.Pn2Error("linear_tc: input has %d channels, layer expects %d" % (cin, layer.cin))
    if rows % pool:
        raise cabi.Pn2Error("linear_tc: rows not divisible by pool")
    if out is None:
        out = torch.empty((rows // pool, layer.cout), dtype=torch.float32, device=x.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows // pool and oc == layer.cout
    rp, ldr = ptr(None), 0
    if res is not None:
        r2, rrows, ldr, rc = _rows2d(res)
        assert rrows == rows and rc == layer.cout
        rp = ptr(r2)
    use_relu = layer.relu if relu is None else relu
    if pool >= 64:
        o2.zero_()      # groups spread over several warps are combined with atomicMax on the output
    cabi.call("pn2_linear_tc_f32", ptr(x2), i32(ldx), ptr(layer.blob), i32(layer.ntile), i32(layer.nchunks),
              i32(layer.nkb), ptr(layer.b), rp, i32(ldr), ptr(o2), i32(ldy), _i64(rows), i32(cin), i32(layer.cout),
              i32(1 if use_relu else 0), i32(pool), work=2.0 * rows * cin * layer.cout)
    return out


def linear_pre(x, cpre, l1, l2, out=None):
    """y = act2(relu(x[:, :cpre] @ W1^T + b1) @ W2^T + b2) in ONE launch (pn2_linear_pre_tc_f32): the tiny first layer
    runs in fp32 inside the operand producers of the second and its output is never materialised.  x (rows, C >= 8) with
    16-byte aligned rows.  Returns None when the shape is not the instantiated one (callers then run the two layers)."""
    x2, rows, ldx, cx = _rows2d(x)
    if (MLP_ENGINE != "tc" or cpre != 5 or l1.cin != cpre or not l1.relu or l2.cin != l1.cout or l2.tc.nchunks != 1
            or cx < 8 or ldx % 4 or x2.data_ptr() % 16):
        return None
    if getattr(l1, "_wpre", None) is None:
        l1._wpre = torch.cat((l1.w[:, :cpre].t().contiguous(), l1.b.view(1, -1)), dim=0).contiguous()    # (cpre + 1, c1)
    tc = l2.tc
    if out is None:
        out = torch.empty((rows, l2.cout), dtype=torch.float32, device=x.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows and oc == l2.cout
    cabi.call("pn2_linear_pre_tc_f32", ptr(x2), i32(ldx), i32(cpre), ptr(l1._wpre), ptr(tc.blob), i32(tc.ntile),
              i32(tc.nchunks), i32(tc.nkb), ptr(tc.b), ptr(o2), i32(ldy), _i64(rows), i32(l2.cin), i32(l2.cout),
              i32(1 if l2.relu else 0), i32(1), work=2.0 * rows * (cpre * l1.cout + l2.cin * l2.cout))
    return out


RCNN_FRONT_FUSED = os.environ.get("PN2_RCNN_FRONT", "1") != "0"


def rcnn_front_supported(x, cpre, off_f, up1, up2, merge, first_f):
    """shape test of pn2_rcnn_front_tc_f32 (csrc/rcnn_front_tc.cu): [5 -> 128 -> 128], [256 -> 128], [128 -> 128]."""
    x2, rows, ldx, cx = _rows2d(x)
    return (RCNN_FRONT_FUSED and MLP_ENGINE == "tc" and cpre == 5 and up1.cin == 5 and up1.cout == 128 and up1.relu
            and up2.cin == 128 and up2.cout == 128 and up2.relu and merge.cin == 256 and merge.cout == 128 and merge.relu
            and first_f.cin == 128 and first_f.cout == 128 and not first_f.relu and off_f >= 8 and off_f % 4 == 0
            and cx >= off_f + 128 and ldx % 4 == 0 and x2.data_ptr() % 16 == 0)


def rcnn_front(x, off_f, up1, up2, merge, first_f, out=None):
    """RCNN input chain in ONE launch: xyz_up_layer [5 -> 128 -> 128] -> merge_down_layer on cat[., rpn features] -> the
    per-point half of SA1's first layer (pre-activation).  x (rows, >= off_f + 128) pooled rows."""
    x2, rows, ldx, _ = _rows2d(x)
    if getattr(up1, "_wpre", None) is None:
        up1._wpre = torch.cat((up1.w[:, :5].t().contiguous(), up1.b.view(1, -1)), dim=0).contiguous()    # (6, 128)
    if out is None:
        out = torch.empty((rows, 128), dtype=torch.float32, device=x.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows and oc == 128
    cabi.call("pn2_rcnn_front_tc_f32", ptr(x2), i32(ldx), i32(off_f), ptr(up1._wpre), ptr(up2.tc.blob), ptr(up2.tc.b),
              ptr(merge.tc.blob), ptr(merge.tc.b), ptr(first_f.tc.blob), ptr(first_f.tc.b), ptr(o2), i32(ldy), _i64(rows),
              work=2.0 * rows * (5 * 128 + 128 * 128 + 256 * 128 + 128 * 128))
    return out


def linear_cat(xa, xb, layer, out=None):
    """y = act(cat[xa, xb] @ W^T + b) without materialising the concatenation on the tensor-core engine
    (pn2_linear_tc2_f32); the exact-fp32 engine concatenates and calls linear()."""
    a2, rows, lda, ca = _rows2d(xa)
    b2, rows_b, ldb, cb = _rows2d(xb)
    if rows != rows_b or ca + cb != layer.cin:
        raise cabi.Pn2Error("linear_cat: shapes do not match the layer")
    aligned = (ca % 64 == 0 and cb % 64 == 0 and lda % 4 == 0 and ldb % 4 == 0 and a2.data_ptr() % 16 == 0
               and b2.data_ptr() % 16 == 0)
    if MLP_ENGINE != "tc" or not aligned:
        return linear(torch.cat((a2, b2), dim=1), layer, out=out)
    tc = layer.tc
    if out is None:
        out = torch.empty((rows, layer.cout), dtype=torch.float32, device=xa.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows and oc == layer.cout
    cabi.call("pn2_linear_tc2_f32", ptr(a2), i32(lda), i32(ca), ptr(b2), i32(ldb), ptr(tc.blob), i32(tc.ntile),
              i32(tc.nchunks), i32(tc.nkb), ptr(tc.b), ptr(None), i32(0), ptr(o2), i32(ldy), _i64(rows), i32(layer.cin),
              i32(layer.cout), i32(1 if layer.relu else 0), i32(1), work=2.0 * rows * layer.cin * layer.cout)
    return out


def sa_group_linear_tc(h, idx, xyz, centres, wxyz, layer, out=None, pool=1):
    B, M, ns = idx.shape
    N = xyz.shape[1]
    h2, hrows, ldh, c1 = _rows2d(h)
    assert hrows == B * N and c1 == layer.cin
    rows = B * M * ns
    if out is None:
        out = torch.empty((rows // pool, layer.cout), dtype=torch.float32, device=h.device)
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == rows // pool and oc == layer.cout
    if pool >= 64:
        o2.zero_()
    cabi.call("pn2_sa_group_linear_tc_f32", ptr(h2), i32(ldh), ptr(idx), ptr(xyz), ptr(centres), ptr(wxyz),
              ptr(layer.blob), i32(layer.ntile), i32(layer.nchunks), i32(layer.nkb), ptr(layer.b), ptr(o2), i32(ldy),
              i32(B), i32(N), i32(M), i32(ns), i32(c1), i32(layer.cout), i32(1 if layer.relu else 0), i32(pool),
              work=2.0 * rows * c1 * (layer.cout + 3))
    return out


def sa_fused_supported(l2, l3, ns):
    """shape test mirroring pn2_sa_fused_tc_f32 (csrc/sa_fused_tc.cu): TMEM columns and shared memory."""
    t2, t3 = l2.tc, l3.tc
    if MLP_ENGINE != "tc" or ns not in (16, 32, 64, 128) or not (l2.relu and l3.relu):
        return False
    if t2.nchunks != 1 or t3.nchunks != 1 or l2.cout % 16 or t2.ntile != l2.cout or 2 * t2.ntile + t3.ntile > 512:
        return False
    smem = (t2.nkb * 2 * t2.ntile * 128 + t3.nkb * 2 * t3.ntile * 128 + 2 * 32768 + 4 * 128 * 16 + 3 * t2.nkb * 64 * 4
            + 2048 + 8192 + 256 + 1024)
    return smem <= 227 * 1024


SA_TRANSPOSED = os.environ.get("PN2_SA_TRANSPOSED", "1") != "0"
# Push only the UNIQUE rows of every ball-query group through the SA MLP (csrc/group_compact.cu): a group that found
# cnt < nsample neighbours is padded with copies of its first hit, which cannot change the max-pool.
SA_SKIP_DUPLICATES = os.environ.get("PN2_SA_SKIP_DUPLICATES", "1") != "0"
# The transposed kernel also handles last layers of fewer than 128 channels (accumulator padded to 128 lanes).  For the two
# RPN SA1 scales (32 / 64 channels, 16 / 32-channel hidden layers) it is no faster than the row-major kernel on DENSE rows
# (0.39 vs 0.43 ms per step: those launches are prologue- and producer-bound), but it has the duplicate-skipping mode, and
# only 8-37 % of those groups' rows are unique: 0.40 ms (row-major, dense) -> 0.24 ms + 0.05 ms of compaction launches,
# +1.5 % scenes/s on the bench (A/B in one call, profiles/r2d_*): routed since round 2.
SA_TRANSPOSED_SMALL = os.environ.get("PN2_SA_TRANSPOSED_SMALL", "1") == "1"
SA_SKIP_MIN_ROWS = 1 << 20      # below ~1 M grouped rows the two compaction launches + the zero-fill cost more than they save


SA_COMPACT_ALIGN = int(os.environ.get("PN2_SA_COMPACT_ALIGN", "8"))


COMPACT_TWO_LAUNCHES = os.environ.get("PN2_COMPACT_TWO_LAUNCHES", "1") != "0"   # 0: count, torch.cumsum, subtraction, lists
COMPACT_USE_HITS = os.environ.get("PN2_COMPACT_USE_HITS", "1") != "0"           # 0: always count the unique rows from the lists


def group_compact(idx, align=None):
    """idx (B, M, ns) int32 from ball_query -> (cmap, jmap int32 lists of the unique rows, device int64 row count);
    no host synchronisation (the count stays on the device).  align: every group is topped up to a multiple of `align`
    rows with copies of its first row (default 8: the SA kernel then pools eight columns at a time)."""
    align = SA_COMPACT_ALIGN if align is None else align
    B, M, ns = idx.shape
    G = B * M
    cnt = torch.empty((G,), dtype=torch.int32, device=idx.device)
    cmap = torch.empty((G * ns + 128,), dtype=torch.int32, device=idx.device)
    jmap = torch.empty((G * ns + 128,), dtype=torch.int32, device=idx.device)
    total = torch.empty((1,), dtype=torch.int64, device=idx.device)
    if COMPACT_TWO_LAUNCHES:
        # count + block sums | offsets + lists: the prefix sum between the two steps happens inside the second kernel
        block_sum = torch.empty(((G + 255) // 256,), dtype=torch.int32, device=idx.device)
        hits = getattr(idx, "_pn2_hits", None) if COMPACT_USE_HITS else None      # the ball query's own hit counts, if it left them
        if hits is not None and (hits.numel() != G or not hits.is_contiguous()):
            hits = None
        cabi.call("pn2_group_compact_lists_i32", ptr(idx), _i64(G), i32(ns), i32(align), ptr(hits), ptr(cnt), ptr(block_sum),
                  ptr(cmap), ptr(jmap), ptr(total))
        return cmap, jmap, total
    cabi.call("pn2_group_unique_count_i32", ptr(idx), _i64(G), i32(ns), i32(align), ptr(cnt))
    incl = torch.cumsum(cnt, dim=0, dtype=torch.int64)
    offs = incl - cnt
    cabi.call("pn2_group_compact_i32", ptr(idx), _i64(G), i32(ns), ptr(cnt), ptr(offs), ptr(cmap), ptr(jmap))
    return cmap, jmap, incl[G - 1:G]


def sa_fused_t_supported(l2, l3, ns):
    """shape test mirroring pn2_sa_fused_t_tc_f32 (csrc/sa_fused_t_tc.cu): last layer of up to 128 channels, or 256."""
    if not SA_TRANSPOSED or MLP_ENGINE != "tc" or ns not in (16, 32, 64, 128) or not (l2.relu and l3.relu):
        return False
    t2 = l2.tc
    if (l3.cout > 128 and l3.cout != 256) or l2.cout % 16 or l2.cout > 128 or t2.nchunks != 1 or t2.ntile != l2.cout:
        return False
    if l3.cout < 128 and not SA_TRANSPOSED_SMALL:
        return False
    nm3 = (l3.cout + 127) // 128
    if l2.cout + nm3 * l2.cout + 128 > 512:          # TMEM: acc2 (x2 when there is room) | W3 | acc3
        return False
    nkb2 = (l2.cout + 63) // 64
    smem = (t2.nkb * 2 * t2.ntile * 128 + nkb2 * 32768 + 2 * 32768 + 4 * 128 * 16 + 3 * t2.nkb * 64 * 4 + 1536 + 2048 + 512 + 1024)
    return smem <= 227 * 1024


def sa_fused_tc(h, idx, xyz, centres, wxyz, l2, l3, out):
    """gather + pair-wise half of layer 1 + layer 2 + layer 3 + max over nsample in one kernel."""
    B, M, ns = idx.shape
    N = xyz.shape[1]
    h2, hrows, ldh, c1 = _rows2d(h)
    assert hrows == B * N and c1 == l2.cin and l3.cin == l2.cout
    o2, orows, ldy, oc = _rows2d(out)
    assert orows == B * M and oc == l3.cout
    t2 = l2.tc
    rows = B * M * ns
    if sa_fused_t_supported(l2, l3, ns):
        w3hi, w3lo = l3.w3t
        cmap = jmap = nrows = None
        if SA_SKIP_DUPLICATES and rows >= SA_SKIP_MIN_ROWS:
            cmap, jmap, nrows = group_compact(idx)
        if ns >= 128 or cmap is not None:
            o2.zero_()
        if cmap is not None and cabi.profiling():
            # bench.py's per-kernel accounting counts the UNIQUE rows (not the copies the 8-row alignment adds back):
            # the conservative figure for roofline.achieved (syncs; profiling only)
            rows = int((idx[:, :, 1:] != idx[:, :, :1]).sum().item()) + B * M
        cabi.call("pn2_sa_fused_t_tc_f32", ptr(h2), i32(ldh), ptr(idx), ptr(xyz), ptr(centres), ptr(wxyz), ptr(t2.blob),
                  i32(t2.ntile), i32(t2.nkb), ptr(t2.b), ptr(w3hi), ptr(w3lo), ptr(l3.b), ptr(o2), i32(ldy), i32(B),
                  i32(N), i32(M), i32(ns), i32(c1), i32(l2.cout), i32(l3.cout), ptr(cmap), ptr(jmap), ptr(nrows),
                  i32(SA_COMPACT_ALIGN if cmap is not None else 1), work=2.0 * rows * (c1 * (l2.cout + 3) + l2.cout * l3.cout))
        return out
    t3 = l3.tc
    if ns >= 64:
        o2.zero_()
    cabi.call("pn2_sa_fused_tc_f32", ptr(h2), i32(ldh), ptr(idx), ptr(xyz), ptr(centres), ptr(wxyz), ptr(t2.blob),
              i32(t2.ntile), i32(t2.nkb), ptr(t2.b), ptr(t3.blob), i32(t3.ntile), i32(t3.nkb), ptr(t3.b), ptr(o2),
              i32(ldy), i32(B), i32(N), i32(M), i32(ns), i32(c1), i32(l2.cout), i32(l3.cout),
              work=2.0 * rows * (c1 * (l2.cout + 3) + l2.cout * l3.cout))
    return out
