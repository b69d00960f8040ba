"""Mirror of pointrcnn/lib/utils/roipool3d/roipool3d_utils.py:7-28 (roipool3d_gpu)."""
import torch

from . import kitti_utils
from . import roipool3d_cuda


def roipool3d_gpu(pts, pts_feature, boxes3d, pool_extra_width, sampled_pt_num=512):
    """pts (B,N,3), pts_feature (B,N,C), boxes3d (B,M,7) -> pooled (B,M,sampled,3+C), empty flag (B,M) int32."""
    batch_size, boxes_num, feature_len = pts.shape[0], boxes3d.shape[1], pts_feature.shape[2]
    pooled_boxes3d = kitti_utils.enlarge_box3d(boxes3d.view(-1, 7), pool_extra_width).view(batch_size, -1, 7)
    pooled_features = torch.zeros((batch_size, boxes_num, sampled_pt_num, 3 + feature_len), dtype=torch.float32,
                                  device=pts.device)
    pooled_empty_flag = torch.zeros((batch_size, boxes_num), dtype=torch.int32, device=pts.device)
    roipool3d_cuda.forward(pts.contiguous(), pooled_boxes3d.contiguous(), pts_feature.contiguous(), pooled_features,
                           pooled_empty_flag)
    return pooled_features, pooled_empty_flag
