"""Mirror of pointrcnn/lib/net/rpn.py: backbone + per-point classification / box-regression
heads + the proposal layer (rpn.py:11-83).  Same attribute names (backbone_net, rpn_cls_layer,
rpn_reg_layer, proposal_layer), Dropout kept at index 1 of each head so state-dict indices
match (rpn.py:26-27,44-45), same init (rpn.py:60-66), same output dict."""
import importlib

import numpy as np
import torch
import torch.nn as nn

from .. import pytorch_utils as pt_utils
from .. import fused as fz
from ..config import cfg
from ..proposal_layer import ProposalLayer


class RPN(pt_utils.PackedCacheMixin, nn.Module):
    def __init__(self, use_xyz=True, mode='TRAIN'):
        super().__init__()
        self.training_mode = (mode == 'TRAIN')
        MODEL = importlib.import_module('.' + cfg.RPN.BACKBONE, package=__package__)
        self.backbone_net = MODEL.get_model(input_channels=int(cfg.RPN.USE_INTENSITY), use_xyz=use_xyz)

        def head(fc_list, out_channels):
            layers, pre = [], cfg.RPN.FP_MLPS[0][-1]
            for c in fc_list:
                layers.append(pt_utils.Conv1d(pre, c, bn=cfg.RPN.USE_BN))
                pre = c
            layers.append(pt_utils.Conv1d(pre, out_channels, activation=None))
            if cfg.RPN.DP_RATIO >= 0:
                layers.insert(1, nn.Dropout(cfg.RPN.DP_RATIO))
            return nn.Sequential(*layers)

        per_loc_bin_num = int(cfg.RPN.LOC_SCOPE / cfg.RPN.LOC_BIN_SIZE) * 2
        reg_channel = per_loc_bin_num * (4 if cfg.RPN.LOC_XZ_FINE else 2) + cfg.RPN.NUM_HEAD_BIN * 2 + 3
        reg_channel += 1  # y offset
        self.rpn_cls_layer = head(cfg.RPN.CLS_FC, 1)
        self.rpn_reg_layer = head(cfg.RPN.REG_FC, reg_channel)
        self.rpn_cls_loss_func = None  # training only (rpn.py:48-56); not part of the inference path
        self.proposal_layer = ProposalLayer(mode=mode)
        self.init_weights()
        self._packed = None

    def init_weights(self):
        if cfg.RPN.LOSS_CLS in ['SigmoidFocalLoss']:
            pi = 0.01
            nn.init.constant_(self.rpn_cls_layer[2].conv.bias, -np.log((1 - pi) / pi))
        nn.init.normal_(self.rpn_reg_layer[-1].conv.weight, mean=0, std=0.001)

    def _source_modules(self):
        return (self.rpn_cls_layer, self.rpn_reg_layer)

    def forward(self, input_data):
        pts_input = input_data['pts_input']
        if self.backbone_net.can_fuse(pts_input):
            backbone_xyz, feats_pm = self.backbone_net.forward_pm(pts_input)          # (B,N,3), (B,N,C)
            if not self._packed_valid():
                cls_l, reg_l = fz.pack_sequential(self.rpn_cls_layer), fz.pack_sequential(self.rpn_reg_layer)
                # the first layers of the two heads read the same (B*N, C) features: one launch with their output channels side
                # by side reads them once (every output element is the same dot product in the same order: bit-identical)
                both = None
                a, b = cls_l[0], reg_l[0]
                if (len(cls_l) > 1 and len(reg_l) > 1 and a.cin == b.cin and a.relu == b.relu and a.cout % 4 == 0
                        and a.cout + b.cout <= 256):
                    both = fz.PackedLayer(torch.cat((a.w[:, :a.cin], b.w[:, :b.cin]), dim=0), torch.cat((a.b, b.b)), a.relu)
                self._store_packed((cls_l, reg_l, both))
            B, N, _ = feats_pm.shape
            cls_l, reg_l, both = self._packed
            x = feats_pm.view(B * N, -1)
            if both is not None:
                h = fz.linear(x, both)                                  # (B*N, c_cls + c_reg)
                heads = ((h[:, :cls_l[0].cout], cls_l[1:]), (h[:, cls_l[0].cout:], reg_l[1:]))
            else:
                heads = ((x, cls_l), (x, reg_l))
            outs = []
            for cur, layers in heads:
                for layer in layers:
                    cur = fz.linear(cur, layer)
                outs.append(cur.view(B, N, -1))
            rpn_cls, rpn_reg = outs
            backbone_features = feats_pm.transpose(1, 2)  # (B,C,N) view, as the reference returns it
        else:
            backbone_xyz, backbone_features = self.backbone_net(pts_input)  # (B,N,3), (B,C,N)
            rpn_cls = self.rpn_cls_layer(backbone_features).transpose(1, 2).contiguous()  # (B,N,1)
            rpn_reg = self.rpn_reg_layer(backbone_features).transpose(1, 2).contiguous()  # (B,N,C)
        return {'rpn_cls': rpn_cls, 'rpn_reg': rpn_reg, 'backbone_xyz': backbone_xyz,
                'backbone_features': backbone_features}
