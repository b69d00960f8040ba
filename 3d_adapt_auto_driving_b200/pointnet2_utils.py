"""Python front of the point-set ops, with the public names and tensor layouts of
pointrcnn/pointnet2_lib/pointnet2/pointnet2_utils.py so that modules written against it keep working:

    furthest_point_sample(xyz (B,N,3), npoint)            -> idx (B,npoint) int32, idx[:, 0] == 0        (:10-36)
    gather_operation(features (B,C,N), idx (B,M))         -> (B,C,M), differentiable in features        (:39-73)
    three_nn(unknown (B,n,3), known (B,m,3))              -> (dist (B,n,3) = sqrt(d2), idx int32)       (:76-105)
    three_interpolate(features (B,C,m), idx, weight)      -> (B,C,n), differentiable in features        (:108-153)
    grouping_operation(features (B,C,N), idx (B,M,S))     -> (B,C,M,S), differentiable in features      (:156-197)
    ball_query(radius, nsample, xyz, new_xyz)             -> idx (B,M,S) int32, zero rows where empty   (:200-228)
    QueryAndGroup, GroupAll                               grouping modules                              (:231-290)

Design: the index-producing ops (sampling, ball query, 3-NN) have no gradient, so they are plain functions; only the
three feature-moving ops are autograd Functions (forward kernel, backward kernel, what backward needs).  Outputs are allocated on the input's device -- the reference's torch.cuda.*Tensor
constructors pin everything to the current CUDA device -- and the work is done by the sm_100a kernels behind
`pointnet2_cuda` (C-ABI).  The inference modules do not come through here when `fused` is on (fused.py reads the
indices straight into the tensor-core SA kernels); this is the op-by-op path and the parity reference for it.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as pointnet2


def _require_contiguous(**tensors):
    for name, t in tensors.items():
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)


def _new(like, shape, dtype, fill=None):
    if fill is None:
        return torch.empty(shape, dtype=dtype, device=like.device)
    return torch.full(shape, fill, dtype=dtype, device=like.device)


# ---- index-producing ops: no gradient flows through them -------------------------------------------------------
@torch.no_grad()
def furthest_point_sample(xyz, npoint):
    _require_contiguous(xyz=xyz)
    b, n = xyz.shape[0], xyz.shape[1]
    idx = _new(xyz, (b, npoint), torch.int32)
    running_min = _new(xyz, (b, n), torch.float32, fill=1e10)       # squared distance to the selected set so far
    pointnet2.furthest_point_sampling_wrapper(b, n, npoint, xyz, running_min, idx)
    return idx


@torch.no_grad()
def ball_query(radius, nsample, xyz, new_xyz):
    _require_contiguous(xyz=xyz, new_xyz=new_xyz)
    b, n, m = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
    idx = _new(xyz, (b, m, nsample), torch.int32, fill=0)           # centres without a neighbour keep the zero row
    pointnet2.ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx)
    return idx


@torch.no_grad()
def three_nn(unknown, known):
    _require_contiguous(unknown=unknown, known=known)
    b, n, m = unknown.shape[0], unknown.shape[1], known.shape[1]
    dist2 = _new(unknown, (b, n, 3), torch.float32)
    idx = _new(unknown, (b, n, 3), torch.int32)
    pointnet2.three_nn_wrapper(b, n, m, unknown, known, dist2, idx)
    return torch.sqrt(dist2), idx                                   # callers get distances, the kernel squared ones


# ---- feature-moving ops: gradient w.r.t. the features only -----------------------------------------------------
class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        _require_contiguous(features=features, idx=idx)
        b, c, n = features.shape
        m = idx.shape[1]
        out = _new(features, (b, c, m), torch.float32)
        pointnet2.gather_points_wrapper(b, c, n, m, features, idx, out)
        ctx.saved = (idx, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        b, c, m = grad_out.shape
        grad = _new(grad_out, (b, c, n), torch.float32, fill=0)
        pointnet2.gather_points_grad_wrapper(b, c, n, m, grad_out.contiguous(), idx, grad)
        return grad, None


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        _require_contiguous(features=features, idx=idx)
        b, c, n = features.shape
        m, s = idx.shape[1], idx.shape[2]
        out = _new(features, (b, c, m, s), torch.float32)
        pointnet2.group_points_wrapper(b, c, n, m, s, features, idx, out)
        ctx.saved = (idx, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        b, c, m, s = grad_out.shape
        grad = _new(grad_out, (b, c, n), torch.float32, fill=0)
        pointnet2.group_points_grad_wrapper(b, c, n, m, s, grad_out.contiguous(), idx, grad)
        return grad, None


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        _require_contiguous(features=features, idx=idx, weight=weight)
        b, c, m = features.shape
        n = idx.shape[1]
        out = _new(features, (b, c, n), torch.float32)
        pointnet2.three_interpolate_wrapper(b, c, m, n, features, idx, weight, out)
        ctx.saved = (idx, weight, m)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        b, c, n = grad_out.shape
        grad = _new(grad_out, (b, c, m), torch.float32, fill=0)
        pointnet2.three_interpolate_grad_wrapper(b, c, n, m, grad_out.contiguous(), idx, weight, grad)
        return grad, None, None


gather_operation = GatherOperation.apply
grouping_operation = GroupingOperation.apply
three_interpolate = ThreeInterpolate.apply


class _ApplyAlias:
    """`FurthestPointSampling.apply(...)`-style access for code written against the reference's Function classes."""

    def __init__(self, fn):
        self.apply = fn


FurthestPointSampling, BallQuery, ThreeNN = _ApplyAlias(furthest_point_sample), _ApplyAlias(ball_query), _ApplyAlias(three_nn)


# ---- grouping modules ----------------------------------------------------------------------------------------------
def _with_xyz(grouped_xyz, grouped_features, use_xyz):
    """Channel layout of a group: [relative xyz (3), features (C)] on dim 1, or one of the two alone."""
    if grouped_features is None:
        if not use_xyz:
            raise ValueError("a group needs features or use_xyz=True")
        return grouped_xyz
    return torch.cat([grouped_xyz, grouped_features], dim=1) if use_xyz else grouped_features


class QueryAndGroup(nn.Module):
    """Ball query around each centre, neighbours' coordinates made relative to it, features gathered alongside."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        rel = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        return _with_xyz(rel, None if features is None else grouping_operation(features, idx), self.use_xyz)   # (B,3+C,M,S)


class GroupAll(nn.Module):
    """One group holding every point; coordinates stay absolute (there is no centre)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        everything = xyz.transpose(1, 2).unsqueeze(2)                                                          # (B,3,1,N)
        if features is None:
            return everything
        return _with_xyz(everything, features.unsqueeze(2), self.use_xyz)
