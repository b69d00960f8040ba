"""Mirror of pointrcnn/lib/rpn/proposal_layer.py: ProposalLayer.forward / distance_based_proposal
/ score_based_proposal.  Decoding is the reference's torch statements (bit-identical boxes).

Two execution paths with identical results:
  * `fused = False`: the reference's control flow -- Python loop over scenes and depth bands,
    boolean-mask indexing (a host synchronisation each), one NMS call per band;
  * default on CUDA: the whole batch at once and without any host synchronisation.  Band
    membership becomes a rank by a cumulative sum along the score-sorted order, the first
    6300 / 2700 members of each band are scattered into fixed-width candidate arrays with
    device-side counts, one batched device NMS per band (pn2_nms_bev_f32, one CTA per scene,
    stops at the post-NMS quota of 70 / 30) and a final scatter assembles the zero-padded
    (B, 100, 7) ROIs.  The reference computes full keep lists on the host and slices them
    (proposal_layer.py:107-112): same first-k result."""
import torch
import torch.nn as nn

from .bbox_transform import decode_bbox_target_torch as decode_bbox_target     # the non-kernel flows: the torch statements
from .config import cfg
from . import kitti_utils
from . import iou3d_utils
from . import glue


class ProposalLayer(nn.Module):
    def __init__(self, mode='TRAIN'):
        super().__init__()
        self.mode = mode
        self.MEAN_SIZE = torch.from_numpy(cfg.CLS_MEAN_SIZE[0])
        if torch.cuda.is_available():
            self.MEAN_SIZE = self.MEAN_SIZE.cuda()
        self.fused = True

    def forward(self, rpn_scores, rpn_reg, xyz):
        """rpn_scores (B,N), rpn_reg (B,N,C), xyz (B,N,3) -> rois (B,M,7), roi scores (B,M)."""
        batch_size = xyz.shape[0]
        batched = (self.fused and rpn_scores.is_cuda and cfg.TEST.RPN_DISTANCE_BASED_PROPOSE
                   and cfg.RPN.NMS_TYPE in ('normal', 'rotate'))
        if batched and glue.ENABLED:
            return self._forward_kernels(rpn_scores, rpn_reg, xyz)
        proposals = decode_bbox_target(xyz.view(-1, 3), rpn_reg.view(-1, rpn_reg.shape[-1]),
                                       anchor_size=self.MEAN_SIZE, loc_scope=cfg.RPN.LOC_SCOPE,
                                       loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                                       get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False)
        proposals[:, 1] += proposals[:, 3] / 2  # y becomes the bottom-face centre
        proposals = proposals.view(batch_size, -1, 7)

        scores = rpn_scores
        _, sorted_idxs = torch.sort(scores, dim=1, descending=True)
        if batched:
            return self._forward_batched(scores, proposals, sorted_idxs)
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

    def _band_sizes(self):
        pre_tot = cfg[self.mode].RPN_PRE_NMS_TOP_N
        post_tot = cfg[self.mode].RPN_POST_NMS_TOP_N
        return ([int(pre_tot * 0.7), pre_tot - int(pre_tot * 0.7)], [int(post_tot * 0.7), post_tot - int(post_tot * 0.7)])

    def _forward_kernels(self, scores, rpn_reg, xyz):
        """The whole layer in six launches (csrc/glue.cu): decode, score order, band selection with the BEV
        boxes of the candidates, one batched device NMS per band, assembly of the zero-padded (B, 100, 7) ROIs.
        Bit-identical to _forward_batched and to the per-scene reference flow (tests/test_glue_gpu.py)."""
        B, N = scores.shape
        pre_n, post_n = self._band_sizes()
        props = glue.decode_bbox(xyz.reshape(-1, 3), rpn_reg.reshape(-1, rpn_reg.shape[-1]), cfg.RPN.LOC_SCOPE,
                                 cfg.RPN.LOC_BIN_SIZE, cfg.RPN.NUM_HEAD_BIN, cfg.CLS_MEAN_SIZE[0],
                                 get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False, y_bottom=True)
        scores = scores.contiguous()
        # one launch instead of torch's eleven-launch segmented radix sort (same order: descending score, index among ties)
        order = glue.argsort_desc(scores) if N <= 16384 else torch.sort(scores, dim=1, descending=True)[1]
        cidx0, cidx1, bev0, bev1, cnt = glue.proposal_select(order, props, pre_n[0], pre_n[1])
        thresh = cfg[self.mode].RPN_NMS_THRESH
        rotated = cfg.RPN.NMS_TYPE == 'rotate'
        # both distance bands in one launch (each keeps only B SMs busy; as two launches the second waited for the first)
        keep0, num0, keep1, num1 = glue.nms_raw_pair(bev0, cnt[0], post_n[0], bev1, cnt[1], post_n[1], thresh, rotated)
        return glue.proposal_assemble(props, scores, cidx0, cidx1, keep0, keep1, num0, num1, post_n[0], post_n[1])

    def _forward_batched(self, scores, proposals, order):
        """distance_based_proposal (proposal_layer.py:58-119) for all scenes, sync-free."""
        from . import iou3d_cuda
        B, N = scores.shape
        dev = scores.device
        pre_tot = cfg[self.mode].RPN_PRE_NMS_TOP_N
        post_tot = cfg[self.mode].RPN_POST_NMS_TOP_N
        pre_n = [int(pre_tot * 0.7), pre_tot - int(pre_tot * 0.7)]
        post_n = [int(post_tot * 0.7), post_tot - int(post_tot * 0.7)]
        thresh = cfg[self.mode].RPN_NMS_THRESH
        rotated = cfg.RPN.NMS_TYPE == 'rotate'

        s_ord = torch.gather(scores, 1, order)                                           # (B,N) descending
        p_ord = torch.gather(proposals, 1, order.unsqueeze(-1).expand(-1, -1, 7))         # (B,N,7)
        dist = p_ord[:, :, 2]
        near = (dist > 0) & (dist <= 40.0)
        far = (dist > 40.0) & (dist <= 80.0)
        near_rank = torch.cumsum(near, dim=1) - 1                                        # rank inside the band
        far_rank = torch.cumsum(far, dim=1) - 1
        far_total = far_rank[:, -1:] + 1
        # an empty far band borrows the near-band candidates that follow the near quota (:92-100)
        borrow = far_total == 0
        far_sel = torch.where(borrow, near & (near_rank >= pre_n[0]), far)
        far_rank = torch.where(borrow, near_rank - pre_n[0], far_rank)

        outs = []
        for sel, rank, pre, post in ((near, near_rank, pre_n[0], post_n[0]), (far_sel, far_rank, pre_n[1], post_n[1])):
            take = sel & (rank < pre)
            cnt = take.sum(dim=1).to(torch.int32)                                        # (B,) on the device
            slot = torch.where(take, rank, torch.full_like(rank, pre))                   # dump slot = pre
            cand = p_ord.new_zeros((B, pre + 1, 7))
            cand.scatter_(1, slot.unsqueeze(-1).expand(-1, -1, 7), p_ord)
            cs = s_ord.new_zeros((B, pre + 1))
            cs.scatter_(1, slot, s_ord)
            cand, cs = cand[:, :pre].contiguous(), cs[:, :pre]
            bev = kitti_utils.boxes3d_to_bev_torch(cand.view(-1, 7)).view(B, pre, 5).contiguous()
            keep, num = iou3d_cuda.nms_device(bev, thresh, rotated=rotated, max_keep=post, counts=cnt)
            outs.append((cand, cs, keep, num.to(torch.int64)))

        ret_bbox3d = scores.new_zeros((B, post_tot + 1, 7))
        ret_scores = scores.new_zeros((B, post_tot + 1))
        base = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        for (cand, cs, keep, num), post in zip(outs, post_n):
            ar = torch.arange(post, device=dev).unsqueeze(0)
            valid = ar < num.unsqueeze(1)
            keep = torch.where(valid, keep, torch.zeros_like(keep))
            kb = torch.gather(cand, 1, keep.unsqueeze(-1).expand(-1, -1, 7))
            ks = torch.gather(cs, 1, keep)
            dst = torch.where(valid, base + ar, torch.full_like(keep, post_tot))          # dump slot = post_tot
            ret_bbox3d.scatter_(1, dst.unsqueeze(-1).expand(-1, -1, 7), kb)
            ret_scores.scatter_(1, dst, ks)
            base = base + num.unsqueeze(1)
        return ret_bbox3d[:, :post_tot].contiguous(), ret_scores[:, :post_tot].contiguous()

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
