"""Mirror of pointrcnn/lib/net/pointnet2_msg.py: the RPN backbone, 4 SA-MSG + 4 FP modules
built from cfg.RPN.SA_CONFIG / cfg.RPN.FP_MLPS (pointnet2_msg.py:7-70), same module names
(SA_modules.k, FP_modules.k) and the same (xyz, features (B,C,N)) return value.

At inference the whole backbone runs on point-major features (forward_pm): no
transpose().contiguous() between layers, MSG/skip concatenations are column slices."""
import torch
import torch.nn as nn

from ..pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG
from ..config import cfg


def get_model(input_channels=6, use_xyz=True):
    return Pointnet2MSG(input_channels=input_channels, use_xyz=use_xyz)


class Pointnet2MSG(nn.Module):
    def __init__(self, input_channels=6, use_xyz=True):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        skip_channel_list = [input_channels]
        sa = cfg.RPN.SA_CONFIG
        for k in range(len(sa.NPOINTS)):
            mlps = [[channel_in] + list(spec) for spec in sa.MLPS[k]]
            channel_out = sum(spec[-1] for spec in mlps)
            self.SA_modules.append(PointnetSAModuleMSG(npoint=sa.NPOINTS[k], radii=sa.RADIUS[k], nsamples=sa.NSAMPLE[k],
                                                       mlps=mlps, use_xyz=use_xyz, bn=cfg.RPN.USE_BN))
            skip_channel_list.append(channel_out)
            channel_in = channel_out
        self.FP_modules = nn.ModuleList()
        fp = cfg.RPN.FP_MLPS
        for k in range(len(fp)):
            pre_channel = fp[k + 1][-1] if k + 1 < len(fp) else channel_out
            self.FP_modules.append(PointnetFPModule(mlp=[pre_channel + skip_channel_list[k]] + list(fp[k])))
        self.fused = True

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward_pm(self, pointcloud):
        """-> xyz (B,N,3), features POINT-major (B,N,C)."""
        xyz = pointcloud[..., 0:3].contiguous()
        feats = pointcloud[..., 3:].contiguous() if pointcloud.size(-1) > 3 else None
        l_xyz, l_feats = [xyz], [feats]
        for level, sa in enumerate(self.SA_modules):
            # from the second level on the cloud being sampled is the previous level's centre list: FPS-ordered
            nx, nf = sa.forward_pm(l_xyz[-1], l_feats[-1], fps_ordered=level > 0)
            l_xyz.append(nx)
            l_feats.append(nf)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_feats[i - 1] = self.FP_modules[i].forward_pm(l_xyz[i - 1], l_xyz[i], l_feats[i - 1], l_feats[i])
        return l_xyz[0], l_feats[0]

    def can_fuse(self, pointcloud):
        return (self.fused and not self.training and pointcloud.is_cuda
                and all(m._can_fuse(pointcloud) for m in self.SA_modules))

    def forward(self, pointcloud):
        if self.can_fuse(pointcloud):
            xyz, feats = self.forward_pm(pointcloud)
            return xyz, feats.transpose(1, 2)  # (B,C,N) view of the point-major result
        xyz, features = self._break_up_pc(pointcloud)
        l_xyz, l_features = [xyz], [features]
        for i in range(len(self.SA_modules)):
            li_xyz, li_features = self.SA_modules[i](l_xyz[i], l_features[i])
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        return l_xyz[0], l_features[0]
