"""Two-stage detector with the interface of pointrcnn/lib/net/point_rcnn.py: PointRCNN(num_classes, use_xyz, mode),
forward(dict) -> dict with the keys eval_rcnn.py reads (point_rcnn.py:26-70).  Sub-module names (`rpn`, `rcnn_net`)
are the state-dict prefixes of the published checkpoints and must not change.

The stages are separate methods (each usable and timeable on its own):
    rpn_stage      per-point scores, box regression and features from the backbone
    proposal_stage foreground mask, depth channel and the distance-banded top-k + NMS proposals (no gradients)
    rcnn_stage     ROI pooling + refinement heads
tests/test_refnet_vs_port_cpu.py / test_refnet_golden_gpu.py pin the data flow against the reference's own forward."""
import torch
import torch.nn as nn

from .rpn import RPN
from .rcnn_net import RCNNNet
from ..config import cfg

RPN_FEATURE_CHANNELS = 128       # width of the backbone's last feature-propagation layer, the RCNN's per-point input


class PointRCNN(nn.Module):
    def __init__(self, num_classes, use_xyz=True, mode='TRAIN'):
        super().__init__()
        if not (cfg.RPN.ENABLED or cfg.RCNN.ENABLED):
            raise ValueError("cfg enables neither the RPN nor the RCNN stage")
        if cfg.RPN.ENABLED:
            self.rpn = RPN(use_xyz=use_xyz, mode=mode)
        if cfg.RCNN.ENABLED:
            if cfg.RCNN.BACKBONE != 'pointnet':
                raise NotImplementedError("RCNN backbone %r (only 'pointnet' is on the inference path)" % cfg.RCNN.BACKBONE)
            self.rcnn_net = RCNNNet(num_classes=num_classes, input_channels=RPN_FEATURE_CHANNELS, use_xyz=use_xyz)

    def rpn_stage(self, batch):
        trainable = self.training and not cfg.RPN.FIXED
        if cfg.RPN.FIXED:
            self.rpn.eval()
        with torch.set_grad_enabled(trainable):
            return self.rpn(batch)

    @torch.no_grad()
    def proposal_stage(self, rpn_out):
        xyz = rpn_out['backbone_xyz']
        scores = rpn_out['rpn_cls'][:, :, 0]
        foreground = (torch.sigmoid(scores) > cfg.RPN.SCORE_THRESH).float()
        rois, roi_scores = self.rpn.proposal_layer(scores, rpn_out['rpn_reg'], xyz)          # (B, M, 7), (B, M)
        results = {'rois': rois, 'roi_scores_raw': roi_scores, 'seg_result': foreground}
        rcnn_in = {'rpn_xyz': xyz, 'rpn_features': rpn_out['backbone_features'].permute((0, 2, 1)), 'seg_mask': foreground,
                   'roi_boxes3d': rois, 'pts_depth': torch.norm(xyz, p=2, dim=2),
                   # for the one-launch RCNN input stage (rcnn_net._pool_rois_canonical): seg_mask is a function of these two
                   'rpn_scores_raw': scores, 'seg_thresh': cfg.RPN.SCORE_THRESH}
        return results, rcnn_in

    def rcnn_stage(self, rcnn_in):
        return self.rcnn_net(rcnn_in)

    def forward(self, input_data):
        if not cfg.RPN.ENABLED:                   # RCNN alone on precomputed RPN outputs (offline mode)
            if not cfg.RCNN.ENABLED:
                raise NotImplementedError
            return self.rcnn_stage(input_data)
        out = dict(self.rpn_stage(input_data))
        if cfg.RCNN.ENABLED:
            results, rcnn_in = self.proposal_stage(out)
            if self.training:
                rcnn_in['gt_boxes3d'] = input_data['gt_boxes3d']
            out.update(results)
            out.update(self.rcnn_stage(rcnn_in))
        return out
