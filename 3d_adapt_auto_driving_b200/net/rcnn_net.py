"""Mirror of pointrcnn/lib/net/rcnn_net.py (inference branch, ROI_SAMPLE_JIT): build the
per-point feature [seg mask, depth/70-0.5, rpn features], pool 512 points per ROI, canonical
transform, xyz_up_layer / merge_down_layer, three SA modules, cls / reg heads
(rcnn_net.py:115-190).  Same attribute names and state-dict layout.  Training branches
(proposal_target_layer, losses) are out of scope of this package."""
import torch
import torch.nn as nn

from ..pointnet2_modules import PointnetSAModule
from .. import pytorch_utils as pt_utils
from .. import fused as fz
from .. import glue
from .. import kitti_utils
from .. import roipool3d_utils
from ..config import cfg


class RCNNNet(pt_utils.PackedCacheMixin, nn.Module):
    def __init__(self, num_classes, input_channels=0, use_xyz=True):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        if cfg.RCNN.USE_RPN_FEATURES:
            self.rcnn_input_channel = 3 + int(cfg.RCNN.USE_INTENSITY) + int(cfg.RCNN.USE_MASK) + int(cfg.RCNN.USE_DEPTH)
            self.xyz_up_layer = pt_utils.SharedMLP([self.rcnn_input_channel] + list(cfg.RCNN.XYZ_UP_LAYER), bn=cfg.RCNN.USE_BN)
            c_out = cfg.RCNN.XYZ_UP_LAYER[-1]
            self.merge_down_layer = pt_utils.SharedMLP([c_out * 2, c_out], bn=cfg.RCNN.USE_BN)
        sa = cfg.RCNN.SA_CONFIG
        for k in range(len(sa.NPOINTS)):
            mlps = [channel_in] + list(sa.MLPS[k])
            npoint = sa.NPOINTS[k] if sa.NPOINTS[k] != -1 else None
            self.SA_modules.append(PointnetSAModule(npoint=npoint, radius=sa.RADIUS[k], nsample=sa.NSAMPLE[k], mlp=mlps,
                                                    use_xyz=use_xyz, bn=cfg.RCNN.USE_BN))
            channel_in = mlps[-1]

        def head(fc_list, out_channels):
            layers, pre = [], channel_in
            for c in fc_list:
                layers.append(pt_utils.Conv1d(pre, c, bn=cfg.RCNN.USE_BN))
                pre = c
            layers.append(pt_utils.Conv1d(pre, out_channels, activation=None))
            if cfg.RCNN.DP_RATIO >= 0:
                layers.insert(1, nn.Dropout(cfg.RCNN.DP_RATIO))
            return nn.Sequential(*layers)

        cls_channel = 1 if num_classes == 2 else num_classes
        self.cls_layer = head(cfg.RCNN.CLS_FC, cls_channel)
        per_loc_bin_num = int(cfg.RCNN.LOC_SCOPE / cfg.RCNN.LOC_BIN_SIZE) * 2
        loc_y_bin_num = int(cfg.RCNN.LOC_Y_SCOPE / cfg.RCNN.LOC_Y_BIN_SIZE) * 2
        reg_channel = per_loc_bin_num * 4 + cfg.RCNN.NUM_HEAD_BIN * 2 + 3
        reg_channel += (1 if not cfg.RCNN.LOC_Y_BY_BIN else loc_y_bin_num * 2)
        self.reg_layer = head(cfg.RCNN.REG_FC, reg_channel)
        self.init_weights(weight_init='xavier')
        self.fused = True
        self._packed = None

    def init_weights(self, weight_init='xavier'):
        init_func = {'kaiming': nn.init.kaiming_normal_, 'xavier': nn.init.xavier_normal_, 'normal': nn.init.normal_}[weight_init]
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d)):
                if weight_init == 'normal':
                    init_func(m.weight, mean=0, std=0.001)
                else:
                    init_func(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.reg_layer[-1].conv.weight, mean=0, std=0.001)

    def _source_modules(self):
        mods = [self.cls_layer, self.reg_layer]
        if cfg.RCNN.USE_RPN_FEATURES:
            mods += [self.xyz_up_layer, self.merge_down_layer]
        return mods

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def _pool_rois(self, input_data):
        """rcnn_net.py:126-154: per-point features, ROI pooling, canonical transform."""
        rpn_xyz, rpn_features = input_data['rpn_xyz'], input_data['rpn_features']
        batch_rois = input_data['roi_boxes3d']
        extra = []
        if cfg.RCNN.USE_INTENSITY:
            extra.append(input_data['rpn_intensity'].unsqueeze(dim=2))
        extra.append(input_data['seg_mask'].unsqueeze(dim=2))
        if cfg.RCNN.USE_DEPTH:
            extra.append((input_data['pts_depth'] / 70.0 - 0.5).unsqueeze(dim=2))
        pts_feature = torch.cat(extra + [rpn_features], dim=2)
        pooled, empty = roipool3d_utils.roipool3d_gpu(rpn_xyz, pts_feature, batch_rois, cfg.RCNN.POOL_EXTRA_WIDTH,
                                                      sampled_pt_num=cfg.RCNN.NUM_POINTS)
        B, M = batch_rois.shape[0], batch_rois.shape[1]
        pooled[:, :, :, 0:3] -= batch_rois[:, :, 0:3].unsqueeze(dim=2)
        flat = pooled.view(B * M, pooled.shape[2], pooled.shape[3])
        kitti_utils.rotate_pc_along_y_torch(flat, batch_rois.view(-1, 7)[:, 6])
        return flat

    _PAD = 8   # column of the rpn feature block in the padded pooled rows (16-byte aligned)

    def _pool_rois_padded(self, input_data):
        """_pool_rois for the fused path: no torch.cat of the per-point features, and pooled rows laid
        out [x y z | mask, depth (| intensity) | 0-pad to column 8 | 128 rpn features] so that the wide
        block is aligned for the tensor-core MLP (pn2_roipool3d_split_f32).  Same sampled points."""
        rpn_xyz, batch_rois = input_data['rpn_xyz'].contiguous(), input_data['roi_boxes3d']
        rpn_features = input_data['rpn_features'].contiguous()
        extra = []
        if cfg.RCNN.USE_INTENSITY:
            extra.append(input_data['rpn_intensity'])
        extra.append(input_data['seg_mask'])
        if cfg.RCNN.USE_DEPTH:
            extra.append(input_data['pts_depth'] / 70.0 - 0.5)
        head = torch.stack(extra, dim=2).contiguous()
        B, N, C2 = rpn_features.shape
        M, S, ld = batch_rois.shape[1], cfg.RCNN.NUM_POINTS, self._PAD + C2
        boxes = kitti_utils.enlarge_box3d(batch_rois.view(-1, 7), cfg.RCNN.POOL_EXTRA_WIDTH).view(B, -1, 7).contiguous()
        pooled = torch.zeros((B, M, S, ld), dtype=torch.float32, device=rpn_xyz.device)
        empty = torch.zeros((B, M), dtype=torch.int32, device=rpn_xyz.device)
        fz.cabi.call("pn2_roipool3d_split_f32", fz.ptr(rpn_xyz), fz.ptr(boxes), fz.ptr(head), fz.i32(head.shape[2]),
                     fz.ptr(rpn_features), fz.i32(C2), fz.ptr(pooled), fz.i32(ld), fz.i32(self._PAD), fz.ptr(empty),
                     fz.i32(B), fz.i32(N), fz.i32(M), fz.i32(S))
        pooled[:, :, :, 0:3] -= batch_rois[:, :, 0:3].unsqueeze(dim=2)
        flat = pooled.view(B * M, S, ld)
        kitti_utils.rotate_pc_along_y_torch(flat, batch_rois.view(-1, 7)[:, 6])
        return flat

    def _pool_rois_canonical(self, input_data):
        """_pool_rois_padded as ONE launch (csrc/roipool3d.cu: pn2_roipool3d_canon_f32) when the per-point extras are
        exactly [seg mask, depth] (default.yaml): enlargement, pooling, mask / depth features from the raw score and the
        point norm, and the canonical transform; the pooled tensor is neither pre-zeroed nor revisited."""
        return glue.roipool_canonical(input_data['rpn_xyz'], input_data['roi_boxes3d'], cfg.RCNN.POOL_EXTRA_WIDTH,
                                      input_data['rpn_scores_raw'], input_data['seg_thresh'], input_data['pts_depth'], 70.0,
                                      input_data['rpn_features'], cfg.RCNN.NUM_POINTS, self._PAD)[0]

    def _fusable(self, probe):
        return (self.fused and not self.training and probe.is_cuda and cfg.RCNN.USE_RPN_FEATURES
                and all(m._can_fuse(probe) for m in self.SA_modules[:-1]))

    def forward(self, input_data):
        if cfg.RCNN.ROI_SAMPLE_JIT:
            if self.training:
                raise NotImplementedError("training (proposal_target_layer) is out of scope of the inference package")
            rf = input_data['rpn_features']
            if (self._fusable(input_data['rpn_xyz']) and rf.shape[2] % 64 == 0
                    and self.rcnn_input_channel <= self._PAD):
                if (glue.ENABLED and 'rpn_scores_raw' in input_data and cfg.RCNN.USE_MASK and cfg.RCNN.USE_DEPTH
                        and not cfg.RCNN.USE_INTENSITY and rf.shape[2] % 4 == 0):
                    return self._forward_fused(self._pool_rois_canonical(input_data), self._PAD)
                return self._forward_fused(self._pool_rois_padded(input_data), self._PAD)
            pts_input = self._pool_rois(input_data)
        else:
            pts_input = input_data['pts_input']

        if self._fusable(pts_input):
            return self._forward_fused(pts_input, self.rcnn_input_channel)

        xyz, features = self._break_up_pc(pts_input)
        if cfg.RCNN.USE_RPN_FEATURES:
            xyz_input = pts_input[..., 0:self.rcnn_input_channel].transpose(1, 2).unsqueeze(dim=3)
            xyz_feature = self.xyz_up_layer(xyz_input)
            rpn_feature = pts_input[..., self.rcnn_input_channel:].transpose(1, 2).unsqueeze(dim=3)
            merged_feature = self.merge_down_layer(torch.cat((xyz_feature, rpn_feature), dim=1))
            l_xyz, l_features = [xyz], [merged_feature.squeeze(dim=3)]
        else:
            l_xyz, l_features = [xyz], [features]
        for i in range(len(self.SA_modules)):
            li_xyz, li_features = self.SA_modules[i](l_xyz[i], l_features[i])
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        rcnn_cls = self.cls_layer(l_features[-1]).transpose(1, 2).contiguous().squeeze(dim=1)
        rcnn_reg = self.reg_layer(l_features[-1]).transpose(1, 2).contiguous().squeeze(dim=1)
        return {'rcnn_cls': rcnn_cls, 'rcnn_reg': rcnn_reg}

    def _forward_fused(self, pts_input, feat_off):
        """pts_input (R, S, ld) point-major rows, R = B * rois: columns [0, rcnn_input_channel) are
        xyz + extras, the rpn features start at column feat_off."""
        if not self._packed_valid():
            up = fz.pack_sequential(self.xyz_up_layer)
            merge = fz.pack_sequential(self.merge_down_layer)[0]
            c = up[-1].cout
            wm = merge.w[:, :merge.cin]
            self._store_packed({
                "up": up, "merge": merge,
                # merge_down on cat[xyz_feature, rpn_feature] = W_a xyz_feature + (W_b rpn_feature + b)
                "merge_a": fz.PackedLayer(wm[:, :c], torch.zeros_like(merge.b), merge.relu),
                "merge_b": fz.PackedLayer(wm[:, c:], merge.b, False),
                "cls": fz.pack_sequential(self.cls_layer), "reg": fz.pack_sequential(self.reg_layer),
            })
        pk = self._packed
        R, S, C = pts_input.shape
        nin = self.rcnn_input_channel
        rows = pts_input.view(R * S, C)
        xyz = pts_input[..., 0:3].contiguous()
        sa1 = self.SA_modules[0]
        sa1_entry = sa1._pack()[0] if sa1._can_fuse(xyz) and len(sa1.groupers) == 1 else None
        first_f = sa1_entry.get("first_f") if sa1_entry else None
        if (len(pk["up"]) == 2 and first_f is not None
                and fz.rcnn_front_supported(rows, nin, feat_off, pk["up"][0], pk["up"][1], pk["merge"], first_f)):
            # xyz_up -> merge_down -> SA1's per-point layer-1 half in one launch: the pooled rows are read once and only
            # H is written (csrc/rcnn_front_tc.cu)
            h = fz.rcnn_front(rows, feat_off, pk["up"][0], pk["up"][1], pk["merge"], first_f)
            l_xyz, l_feats = sa1.forward_pm(xyz, None, h_first=h)
            levels = list(enumerate(self.SA_modules))[1:]
        else:
            cur = None
            if len(pk["up"]) == 2:
                cur = fz.linear_pre(rows, nin, pk["up"][0], pk["up"][1])   # [5 -> 128 -> 128] in one launch, layer 1 on the fly
            if cur is None:
                cur = rows[:, 0:nin]
                for layer in pk["up"]:
                    cur = fz.linear(cur, layer)
            rpn_feat = rows[:, feat_off:]
            if fz.MLP_ENGINE == "tc" and feat_off % 4 == 0 and C % 4 == 0 and rpn_feat.shape[1] % 64 == 0:
                merged = fz.linear_cat(cur, rpn_feat, pk["merge"])            # one GEMM over the virtual concatenation
            else:
                side = fz.linear(rpn_feat, pk["merge_b"])                     # W_b rpn_feature + b
                merged = fz.linear(cur, pk["merge_a"], res=side)              # relu(W_a xyz_feature + side)
            l_xyz, l_feats = xyz, merged.view(R, S, -1)
            levels = list(enumerate(self.SA_modules))
        for level, sa in levels:
            l_xyz, l_feats = sa.forward_pm(l_xyz, l_feats, fps_ordered=level > 0)
        feat = l_feats.reshape(R, -1)
        outs = []
        for layers in (pk["cls"], pk["reg"]):
            cur = feat
            for layer in layers:
                cur = fz.linear(cur, layer)
            outs.append(cur)
        return {'rcnn_cls': outs[0], 'rcnn_reg': outs[1]}
