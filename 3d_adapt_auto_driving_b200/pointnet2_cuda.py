"""Drop-in for the reference's `pointnet2_cuda` extension module.

Same nine functions, same positional arguments and caller-allocated outputs as
pointrcnn/pointnet2_lib/pointnet2/src/pointnet2_api.cpp:10-24, so code written against the
reference extension (pointnet2_utils.py:7 `import pointnet2_cuda as pointnet2`) keeps working.
Each one forwards raw device pointers to the C-ABI (include/pn2_b200.h) on torch's current
stream.  Unlike the reference wrappers (only ball_query.cpp:16-17 checks anything) dtype,
device and contiguity violations raise here instead of corrupting memory.
"""
import torch

from . import cabi
from .cabi import i32, f32, ptr


def _chk(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise cabi.Pn2Error("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise cabi.Pn2Error("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise cabi.Pn2Error("%s must be contiguous" % name)
    return t


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _chk(points, torch.float32, "points"); _chk(temp, torch.float32, "temp"); _chk(idx, torch.int32, "idx")
    cabi.call("pn2_fps_f32", ptr(points), ptr(temp), ptr(idx), i32(b), i32(n), i32(m))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx"); _chk(out, torch.float32, "out")
    cabi.call("pn2_gather_points_f32", ptr(points), ptr(idx), ptr(out), i32(b), i32(c), i32(n), i32(npoints))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx"); _chk(grad_points, torch.float32, "grad_points")
    cabi.call("pn2_gather_points_grad_f32", ptr(grad_out), ptr(idx), ptr(grad_points), i32(b), i32(c), i32(n), i32(npoints))
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _chk(new_xyz, torch.float32, "new_xyz"); _chk(xyz, torch.float32, "xyz"); _chk(idx, torch.int32, "idx")
    # same idx through the culled / ballot-append scan (ball_query.cu); it needs b*m int32 of scratch
    order = torch.empty((b, m), dtype=torch.int32, device=xyz.device) if n >= 128 else None
    cabi.call("pn2_ball_query_culled_f32", ptr(new_xyz), ptr(xyz), ptr(idx), ptr(None), ptr(order), i32(b), i32(n), i32(m),
              f32(radius), i32(nsample), f32(0.0), i32(0), work=12.0 * b * m * n)
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx"); _chk(out, torch.float32, "out")
    cabi.call("pn2_group_points_f32", ptr(points), ptr(idx), ptr(out), i32(b), i32(c), i32(n), i32(npoints), i32(nsample))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx"); _chk(grad_points, torch.float32, "grad_points")
    cabi.call("pn2_group_points_grad_f32", ptr(grad_out), ptr(idx), ptr(grad_points), i32(b), i32(c), i32(n), i32(npoints), i32(nsample))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _chk(unknown, torch.float32, "unknown"); _chk(known, torch.float32, "known")
    _chk(dist2, torch.float32, "dist2"); _chk(idx, torch.int32, "idx")
    # large levels take the spatially culled scan (same outputs); it needs b*n int32 of scratch
    order = torch.empty((b, n), dtype=torch.int32, device=unknown.device) if (n >= 1024 and m >= 512) else None
    cabi.call("pn2_three_nn_culled_f32", ptr(unknown), ptr(known), ptr(dist2), ptr(idx), ptr(order), i32(b), i32(n),
              i32(m), work=12.0 * b * n * m)      # scan bytes: 12 B per unknown-known pair


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
    _chk(weight, torch.float32, "weight"); _chk(out, torch.float32, "out")
    cabi.call("pn2_three_interpolate_f32", ptr(points), ptr(idx), ptr(weight), ptr(out), i32(b), i32(c), i32(m), i32(n))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
    _chk(weight, torch.float32, "weight"); _chk(grad_points, torch.float32, "grad_points")
    cabi.call("pn2_three_interpolate_grad_f32", ptr(grad_out), ptr(idx), ptr(weight), ptr(grad_points), i32(b), i32(c), i32(n), i32(m))
