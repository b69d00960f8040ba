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
    # Replay the whole inference forward as ONE CUDA graph per input shape (forward() below).  Off by default: callers that
    # toggle kernel-selection flags between calls, or capture their own graph around the model (inference.Detector), want
    # the eager launches.  The drop-in tree of the unmodified eval_rcnn.py switches it on (evaltree.py): that script drives
    # the model from one Python thread, where the ~170 eager launches of a batch cost more host time than the GPU needs
    # to execute them.
    graph_forward = False

    def __init__(self, num_classes, use_xyz=True, mode='TRAIN'):
        super().__init__()
        self._graphs = {}
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

    # ---- whole-forward CUDA graph (opt-in, see graph_forward) ----
    def train(self, mode=True):
        self._graphs.clear()                      # new mode / new weights: captured graphs are stale
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self._graphs.clear()
        return super().load_state_dict(*args, **kwargs)

    def _graphable(self, input_data):
        pts = input_data.get('pts_input') if isinstance(input_data, dict) else None
        return (self.graph_forward and not self.training and not torch.is_grad_enabled() and cfg.RPN.ENABLED and cfg.RCNN.ENABLED
                and isinstance(pts, torch.Tensor) and pts.is_cuda and pts.dtype == torch.float32 and pts.dim() == 3
                and set(input_data) == {'pts_input'} and not torch.cuda.is_current_stream_capturing())

    def _forward_graphed(self, input_data):
        pts = input_data['pts_input']
        key = (tuple(pts.shape), pts.device)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 2:            # e.g. the last, smaller batch of an epoch: keep two shapes at most
                self._graphs.pop(next(iter(self._graphs)))
            static_in = pts.clone()
            side = torch.cuda.Stream(device=pts.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):         # warm-up off the capture: lazy weight packing, workspaces
                for _ in range(2):
                    self._forward_eager({'pts_input': static_in})
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_out = self._forward_eager({'pts_input': static_in})
            entry = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(pts, non_blocking=True)
        graph.replay()
        # the static buffers are rewritten by the next replay: hand out copies (0.3 GB/s-scale, ~0.1 ms per batch)
        return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in static_out.items()}

    def forward(self, input_data):
        if self._graphable(input_data):
            return self._forward_graphed(input_data)
        return self._forward_eager(input_data)

    def _forward_eager(self, input_data):
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
