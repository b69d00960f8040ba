"""CPU port of the reference's PointRCNN inference composition.  TEST / BASELINE INFRASTRUCTURE
ONLY (tests/, smoke(), bench.py cpu_baseline and --impl reference); the product never imports it.

The reference has no CPU implementation of this path (pointnet2_utils.py:25-26 hard-codes
torch.cuda tensors), so this is a PORT: every native op is the C restatement of the
reference kernel (oracle/pn2_oracle.c, geom_oracle.c), every 1x1 conv / BN / ReLU / max-pool
is the model's own torch module evaluated on the CPU in fp32, and the order of operations is
the reference's:
  PointnetSAModuleMSG.forward   pointnet2_modules.py:19-55
  QueryAndGroup.forward         pointnet2_utils.py:241-264
  PointnetFPModule.forward      pointnet2_modules.py:127-156
  Pointnet2MSG.forward          lib/net/pointnet2_msg.py:56-70
  RPN.forward                   lib/net/rpn.py:68-83
  ProposalLayer.forward         lib/rpn/proposal_layer.py:15-141
  RCNNNet.forward               lib/net/rcnn_net.py:115-190
  eval post-processing          tools/eval_rcnn.py:516-535, 611-627
It takes a model built by the product package (same parameters) moved to the CPU.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def query_and_group(grouper, xyz, new_xyz, features):
    """pointnet2_utils.py:241-264 -> (B, 3+C, M, ns)."""
    idx = orc.ball_query(grouper.radius, grouper.nsample, xyz.numpy(), new_xyz.numpy())
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = _t(orc.group_points(xyz_trans.numpy(), idx))
    grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
    if features is None:
        return grouped_xyz
    grouped = _t(orc.group_points(features.contiguous().numpy(), idx))
    return torch.cat([grouped_xyz, grouped], dim=1) if grouper.use_xyz else grouped


def group_all(grouper, xyz, features):
    """pointnet2_utils.py:267-290."""
    grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
    if features is None:
        return grouped_xyz
    grouped = features.unsqueeze(2)
    return torch.cat([grouped_xyz, grouped], dim=1) if grouper.use_xyz else grouped


def sa_forward(sa, xyz, features):
    """pointnet2_modules.py:19-55."""
    new_xyz = None
    if sa.npoint is not None:
        idx, _ = orc.fps(xyz.numpy(), sa.npoint)
        new_xyz = _t(orc.gather_points(xyz.transpose(1, 2).contiguous().numpy(), idx)).transpose(1, 2).contiguous()
    outs = []
    for grouper, mlp in zip(sa.groupers, sa.mlps):
        if sa.npoint is not None:
            g = query_and_group(grouper, xyz, new_xyz, features)
        else:
            g = group_all(grouper, xyz, features)
        g = mlp(g)
        g = F.max_pool2d(g, kernel_size=[1, g.size(3)])
        outs.append(g.squeeze(-1))
    return new_xyz, torch.cat(outs, dim=1)


def fp_forward(fp, unknown, known, unknow_feats, known_feats):
    """pointnet2_modules.py:127-156."""
    d2, idx = orc.three_nn(unknown.numpy(), known.numpy())
    dist = torch.sqrt(_t(d2))                         # pointnet2_utils.py:98
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    interp = _t(orc.three_interpolate(known_feats.contiguous().numpy(), idx, weight.numpy()))
    new = torch.cat([interp, unknow_feats], dim=1) if unknow_feats is not None else interp
    return fp.mlp(new.unsqueeze(-1)).squeeze(-1)


def backbone_forward(net, pointcloud):
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    l_xyz, l_features = [xyz], [features]
    for sa in net.SA_modules:
        nx, nf = sa_forward(sa, l_xyz[-1], l_features[-1])
        l_xyz.append(nx)
        l_features.append(nf)
    for i in range(-1, -(len(net.FP_modules) + 1), -1):
        l_features[i - 1] = fp_forward(net.FP_modules[i], l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
    return l_xyz[0], l_features[0]


def rpn_forward(rpn, pts_input):
    xyz, feats = backbone_forward(rpn.backbone_net, pts_input)
    rpn_cls = rpn.rpn_cls_layer(feats).transpose(1, 2).contiguous()
    rpn_reg = rpn.rpn_reg_layer(feats).transpose(1, 2).contiguous()
    return {'rpn_cls': rpn_cls, 'rpn_reg': rpn_reg, 'backbone_xyz': xyz, 'backbone_features': feats}


def _boxes3d_to_bev(b):
    cu, cv, hl, hw = b[:, 0], b[:, 2], b[:, 5] / 2, b[:, 4] / 2
    return torch.stack((cu - hl, cv - hw, cu + hl, cv + hw, b[:, 6]), dim=1)


def _nms(bev, scores, thresh, rotated):
    """iou3d_utils.py:56-87: sort by score, NMS, indices into the original order."""
    order = scores.sort(0, descending=True)[1]
    b = bev[order].contiguous().numpy()
    keep = orc.nms_rotated(b, thresh) if rotated else orc.nms_normal(b, thresh)
    return order[torch.from_numpy(keep)]


def proposal_forward(pkg, rpn_scores, rpn_reg, xyz, mode='TEST'):
    """lib/rpn/proposal_layer.py:15-141 (distance-based, NMS_TYPE per cfg)."""
    cfg = pkg["cfg"]
    decode = pkg["decode_bbox_target"]
    B = xyz.shape[0]
    mean_size = torch.from_numpy(cfg.CLS_MEAN_SIZE[0])
    props = decode(xyz.reshape(-1, 3), rpn_reg.reshape(-1, rpn_reg.shape[-1]), anchor_size=mean_size,
                   loc_scope=cfg.RPN.LOC_SCOPE, loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                   get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False)
    props[:, 1] += props[:, 3] / 2
    props = props.view(B, -1, 7)
    order_all = torch.sort(rpn_scores, dim=1, descending=True)[1]
    top_n = cfg[mode].RPN_POST_NMS_TOP_N
    ret_b = rpn_scores.new_zeros((B, top_n, 7))
    ret_s = rpn_scores.new_zeros((B, top_n))
    pre_tot = cfg[mode].RPN_PRE_NMS_TOP_N
    pre_n = [0, int(pre_tot * 0.7), pre_tot - int(pre_tot * 0.7)]
    post_n = [0, int(top_n * 0.7), top_n - int(top_n * 0.7)]
    edges = [0, 40.0, 80.0]
    rotated = cfg.RPN.NMS_TYPE == 'rotate'
    for k in range(B):
        so, po = rpn_scores[k][order_all[k]], props[k][order_all[k]]
        dist = po[:, 2]
        first = (dist > edges[0]) & (dist <= edges[1])
        out_s, out_p = [], []
        for i in range(1, 3):
            band = (dist > edges[i - 1]) & (dist <= edges[i])
            if band.sum() != 0:
                cs, cp = so[band][:pre_n[i]], po[band][:pre_n[i]]
            else:
                cs, cp = so[first][pre_n[i - 1]:][:pre_n[i]], po[first][pre_n[i - 1]:][:pre_n[i]]
            keep = _nms(_boxes3d_to_bev(cp), cs, cfg[mode].RPN_NMS_THRESH, rotated)[:post_n[i]]
            out_s.append(cs[keep]); out_p.append(cp[keep])
        s, p = torch.cat(out_s), torch.cat(out_p)
        ret_b[k, :p.size(0)] = p
        ret_s[k, :s.size(0)] = s
    return ret_b, ret_s


def rcnn_forward(pkg, rcnn, info):
    """lib/net/rcnn_net.py:115-190 (eval, ROI_SAMPLE_JIT)."""
    cfg = pkg["cfg"]
    xyz, feats, rois = info['rpn_xyz'], info['rpn_features'], info['roi_boxes3d']
    extra = [info['seg_mask'].unsqueeze(2)]
    if cfg.RCNN.USE_DEPTH:
        extra.append((info['pts_depth'] / 70.0 - 0.5).unsqueeze(2))
    pts_feature = torch.cat(extra + [feats], dim=2)
    B, M = rois.shape[0], rois.shape[1]
    big = rois.reshape(-1, 7).clone()
    big[:, 3:6] += cfg.RCNN.POOL_EXTRA_WIDTH * 2
    big[:, 1] += cfg.RCNN.POOL_EXTRA_WIDTH
    pooled, _ = orc.roipool3d(xyz.numpy(), pts_feature.numpy(), big.view(B, M, 7).numpy(), sampled=cfg.RCNN.NUM_POINTS)
    pooled = _t(pooled)
    pooled[:, :, :, 0:3] -= rois[:, :, 0:3].unsqueeze(2)
    flat = pooled.view(B * M, pooled.shape[2], pooled.shape[3])
    ry = rois.reshape(-1, 7)[:, 6]
    # kitti_utils.py:45-63: [x', z'] = [x, z] @ [[cos, -sin], [sin, cos]]^T as a batched matmul (its accumulation,
    # not x*cos - z*sin written out: the two differ in the last bit)
    cosa, sina = torch.cos(ry).view(-1, 1), torch.sin(ry).view(-1, 1)
    rot = torch.stack((torch.cat((cosa, -sina), dim=1), torch.cat((sina, cosa), dim=1)), dim=1)      # (B*M, 2, 2)
    xz = torch.matmul(torch.stack((flat[:, :, 0], flat[:, :, 2]), dim=2), rot.permute(0, 2, 1))
    flat[:, :, 0] = xz[:, :, 0]
    flat[:, :, 2] = xz[:, :, 1]
    nin = rcnn.rcnn_input_channel
    pxyz = flat[..., 0:3].contiguous()
    xyz_feature = rcnn.xyz_up_layer(flat[..., 0:nin].transpose(1, 2).unsqueeze(3))
    rpn_feature = flat[..., nin:].transpose(1, 2).unsqueeze(3)
    merged = rcnn.merge_down_layer(torch.cat((xyz_feature, rpn_feature), dim=1))
    l_xyz, l_feat = pxyz, merged.squeeze(3)
    for sa in rcnn.SA_modules:
        l_xyz, l_feat = sa_forward(sa, l_xyz, l_feat)
    cls = rcnn.cls_layer(l_feat).transpose(1, 2).contiguous().squeeze(1)
    reg = rcnn.reg_layer(l_feat).transpose(1, 2).contiguous().squeeze(1)
    return {'rcnn_cls': cls, 'rcnn_reg': reg}


def pointrcnn_forward(pkg, model, pts_input):
    """lib/net/point_rcnn.py:26-70 on the CPU.  pkg = {"cfg", "decode_bbox_target"} from the product."""
    cfg = pkg["cfg"]
    with torch.no_grad():
        out = rpn_forward(model.rpn, pts_input)
        scores_raw = out['rpn_cls'][:, :, 0]
        seg_mask = (torch.sigmoid(scores_raw) > cfg.RPN.SCORE_THRESH).float()
        depth = torch.norm(out['backbone_xyz'], p=2, dim=2)
        rois, roi_scores = proposal_forward(pkg, scores_raw, out['rpn_reg'], out['backbone_xyz'])
        out.update({'rois': rois, 'roi_scores_raw': roi_scores, 'seg_result': seg_mask})
        out.update(rcnn_forward(pkg, model.rcnn_net, {
            'rpn_xyz': out['backbone_xyz'], 'rpn_features': out['backbone_features'].permute(0, 2, 1),
            'seg_mask': seg_mask, 'roi_boxes3d': rois, 'pts_depth': depth}))
    return out


def postprocess(pkg, out, batch_size):
    """tools/eval_rcnn.py:516-535, 611-627 -> per scene (boxes3d (k,7), raw scores (k,))."""
    cfg = pkg["cfg"]
    decode = pkg["decode_bbox_target"]
    rois = out['rois']
    cls = out['rcnn_cls'].view(batch_size, -1, out['rcnn_cls'].shape[1])
    reg = out['rcnn_reg'].view(batch_size, -1, out['rcnn_reg'].shape[1])
    pred = decode(rois.reshape(-1, 7), reg.reshape(-1, reg.shape[-1]), anchor_size=torch.from_numpy(cfg.CLS_MEAN_SIZE[0]),
                  loc_scope=cfg.RCNN.LOC_SCOPE, loc_bin_size=cfg.RCNN.LOC_BIN_SIZE, num_head_bin=cfg.RCNN.NUM_HEAD_BIN,
                  get_xz_fine=True, get_y_by_bin=cfg.RCNN.LOC_Y_BY_BIN, loc_y_scope=cfg.RCNN.LOC_Y_SCOPE,
                  loc_y_bin_size=cfg.RCNN.LOC_Y_BIN_SIZE, get_ry_fine=True).view(batch_size, -1, 7)
    raw = cls
    inds = torch.sigmoid(raw) > cfg.RCNN.SCORE_THRESH
    res = []
    for k in range(batch_size):
        cur = inds[k].view(-1)
        if cur.sum() == 0:
            res.append((np.zeros((0, 7), np.float32), np.zeros((0,), np.float32)))
            continue
        b, s = pred[k, cur], raw[k, cur].view(-1)
        keep = _nms(_boxes3d_to_bev(b), s, cfg.RCNN.NMS_THRESH, True)
        res.append((b[keep].numpy(), s[keep].numpy()))
    return res
