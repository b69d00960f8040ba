"""Mirror of pointrcnn/lib/rpn/proposal_layer.py: ProposalLayer.forward / distance_based_proposal
/ score_based_proposal.  Decoding and band selection are the reference's torch statements; the
NMS runs on the device and stops after the post-NMS quota (the reference computes the full
keep list on the host and then slices it, proposal_layer.py:107-112 -- same first-k result)."""
import torch
import torch.nn as nn

from .bbox_transform import decode_bbox_target
from .config import cfg
from . import kitti_utils
from . import iou3d_utils


class ProposalLayer(nn.Module):
    def __init__(self, mode='TRAIN'):
        super().__init__()
        self.mode = mode
        self.MEAN_SIZE = torch.from_numpy(cfg.CLS_MEAN_SIZE[0])
        if torch.cuda.is_available():
            self.MEAN_SIZE = self.MEAN_SIZE.cuda()

    def forward(self, rpn_scores, rpn_reg, xyz):
        """rpn_scores (B,N), rpn_reg (B,N,C), xyz (B,N,3) -> rois (B,M,7), roi scores (B,M)."""
        batch_size = xyz.shape[0]
        proposals = decode_bbox_target(xyz.view(-1, 3), rpn_reg.view(-1, rpn_reg.shape[-1]),
                                       anchor_size=self.MEAN_SIZE, loc_scope=cfg.RPN.LOC_SCOPE,
                                       loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                                       get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False)
        proposals[:, 1] += proposals[:, 3] / 2  # y becomes the bottom-face centre
        proposals = proposals.view(batch_size, -1, 7)

        scores = rpn_scores
        _, sorted_idxs = torch.sort(scores, dim=1, descending=True)
        top_n = cfg[self.mode].RPN_POST_NMS_TOP_N
        ret_bbox3d = scores.new_zeros((batch_size, top_n, 7))
        ret_scores = scores.new_zeros((batch_size, top_n))
        for k in range(batch_size):
            if cfg.TEST.RPN_DISTANCE_BASED_PROPOSE:
                s, p = self.distance_based_proposal(scores[k], proposals[k], sorted_idxs[k])
            else:
                s, p = self.score_based_proposal(scores[k], proposals[k], sorted_idxs[k])
            tot = p.size(0)
            ret_bbox3d[k, :tot] = p
            ret_scores[k, :tot] = s
        return ret_bbox3d, ret_scores

    def _nms(self, boxes_bev, scores, keep_n):
        thresh = cfg[self.mode].RPN_NMS_THRESH
        if cfg.RPN.NMS_TYPE == 'rotate':
            return iou3d_utils.nms_gpu(boxes_bev, scores, thresh, max_keep=keep_n)
        if cfg.RPN.NMS_TYPE == 'normal':
            return iou3d_utils.nms_normal_gpu(boxes_bev, scores, thresh, max_keep=keep_n)
        raise NotImplementedError

    def distance_based_proposal(self, scores, proposals, order):
        """two depth bands (0,40] and (40,80] on the decoded z, 70 % / 30 % of the pre- and
        post-NMS budgets; an empty far band borrows the next candidates of the near band."""
        edges = [0, 40.0, 80.0]
        pre_tot = cfg[self.mode].RPN_PRE_NMS_TOP_N
        pre_n = [0, int(pre_tot * 0.7), pre_tot - int(pre_tot * 0.7)]
        post_tot = cfg[self.mode].RPN_POST_NMS_TOP_N
        post_n = [0, int(post_tot * 0.7), post_tot - int(post_tot * 0.7)]

        scores_ordered = scores[order]
        proposals_ordered = proposals[order]
        dist = proposals_ordered[:, 2]
        first_mask = (dist > edges[0]) & (dist <= edges[1])
        out_s, out_p = [], []
        for i in range(1, len(edges)):
            band = (dist > edges[i - 1]) & (dist <= edges[i])
            if band.sum() != 0:
                cur_scores = scores_ordered[band][:pre_n[i]]
                cur_proposals = proposals_ordered[band][:pre_n[i]]
            else:
                assert i == 2, '%d' % i
                cur_scores = scores_ordered[first_mask][pre_n[i - 1]:][:pre_n[i]]
                cur_proposals = proposals_ordered[first_mask][pre_n[i - 1]:][:pre_n[i]]
            boxes_bev = kitti_utils.boxes3d_to_bev_torch(cur_proposals)
            keep_idx = self._nms(boxes_bev, cur_scores, post_n[i])[:post_n[i]]
            out_s.append(cur_scores[keep_idx])
            out_p.append(cur_proposals[keep_idx])
        return torch.cat(out_s, dim=0), torch.cat(out_p, dim=0)

    def score_based_proposal(self, scores, proposals, order):
        pre = cfg[self.mode].RPN_PRE_NMS_TOP_N
        post = cfg[self.mode].RPN_POST_NMS_TOP_N
        cur_scores = scores[order][:pre]
        cur_proposals = proposals[order][:pre]
        boxes_bev = kitti_utils.boxes3d_to_bev_torch(cur_proposals)
        keep_idx = iou3d_utils.nms_gpu(boxes_bev, cur_scores, cfg[self.mode].RPN_NMS_THRESH, max_keep=post)[:post]
        return cur_scores[keep_idx], cur_proposals[keep_idx]
