"""Drop-in for the reference's `roipool3d_cuda.forward`
(pointrcnn/lib/utils/roipool3d/src/roipool3d.cpp:17-45, :198-203)."""
import torch

from . import cabi
from .cabi import i32, ptr


def forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    for t, name in ((xyz, "xyz"), (boxes3d, "boxes3d"), (pts_feature, "pts_feature"), (pooled_features, "pooled_features")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise cabi.Pn2Error("%s must be a contiguous CUDA float32 tensor" % name)
    if not (pooled_empty_flag.is_cuda and pooled_empty_flag.dtype == torch.int32):
        raise cabi.Pn2Error("pooled_empty_flag must be a CUDA int32 tensor")
    B, N, _ = xyz.shape
    M = boxes3d.size(1)
    C = pts_feature.size(2)
    S = pooled_features.size(2)
    cabi.call("pn2_roipool3d_f32", ptr(xyz), ptr(boxes3d), ptr(pts_feature), ptr(pooled_features), ptr(pooled_empty_flag),
              i32(B), i32(N), i32(M), i32(C), i32(S))
    return 1
