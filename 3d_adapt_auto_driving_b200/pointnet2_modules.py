"""Mirror of pointrcnn/pointnet2_lib/pointnet2/pointnet2_modules.py: PointnetSAModuleMSG,
PointnetSAModule, PointnetFPModule with the reference's keyword-only constructors, forward
signatures, return layouts and parameter names.

Two execution paths share the same parameters:
  * reference-structured (training, CPU-free fallback never: it still needs the CUDA ops):
    furthest_point_sample -> gather -> QueryAndGroup -> SharedMLP (torch conv) -> max_pool2d,
    i.e. the op-level API exactly as the reference composes it.  This is also the plain fp32
    PyTorch reference the tests compare the fused path against.
  * fused inference (`module.eval()` on CUDA, default): forward_pm() on POINT-major features;
    per scale: per-point half of layer 1 -> [gather + xyz half of layer 1 + layer 2] in one
    kernel -> [layer 3 + max over nsample] in one kernel, BatchNorm folded, both MSG scales
    written straight into their column slice of the output.  Set `module.fused = False` to
    force the reference-structured path.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils
from . import fused as fz


class _PointnetSAModuleBase(pt_utils.PackedCacheMixin, nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'
        self.fused = True
        self._packed = None

    # ---- cache of folded weights for the fused path (pt_utils.PackedCacheMixin) ----
    def _pack(self):
        if not self._packed_valid():
            packed = []
            for mlp in self.mlps:
                layers = fz.pack_sequential(mlp)
                first = layers[0]
                w1 = first.w[:, :first.cin]
                entry = {"layers": layers}
                if isinstance(self.groupers[0], pointnet2_utils.QueryAndGroup):
                    entry["wxyz"] = w1[:, :3].t().contiguous()                       # (3, c1)
                    entry["first_f"] = fz.PackedLayer(w1[:, 3:], first.b, False) if first.cin > 3 else None
                    entry["b1"] = first.b
                packed.append(entry)
            self._store_packed(packed)
        return self._packed

    def _can_fuse(self, xyz):
        if not (self.fused and not self.training and xyz.is_cuda and self.pool_method == 'max_pool'):
            return False
        if any(not getattr(g, "use_xyz", True) for g in self.groupers):
            return False
        if any(len(list(m.children())) < 2 or not pt_utils.foldable(m) for m in self.mlps):
            return False
        for g in self.groupers:
            if isinstance(g, pointnet2_utils.QueryAndGroup) and (128 % g.nsample or g.nsample % 4):
                return False
            if isinstance(g, pointnet2_utils.GroupAll) and (128 % xyz.shape[1] or xyz.shape[1] % 4):
                return False      # the pooled-linear kernel needs the group size to divide its 128-row tile
        return True

    def forward_pm(self, xyz, feats_pm=None, new_xyz=None, fps_ordered=False, h_first=None):
        """xyz (B,N,3), feats_pm (B,N,C) point-major or None -> (new_xyz (B,M,3) or None,
        out (B,M,sum C_out) point-major).  fps_ordered: xyz is itself the centre list of a previous FPS (fused.fps_gather).
        h_first: (B*N, c1) the per-point half of the first layer already computed by the caller (single-scale modules;
        RCNNNet fuses it into its input chain) -- feats_pm is then not read."""
        B, N, _ = xyz.shape
        packed = self._pack()
        group_all = isinstance(self.groupers[0], pointnet2_utils.GroupAll)
        if group_all:
            if 128 % N or N % 4:
                raise NotImplementedError("GroupAll fused path needs N to divide 128")
            x = xyz if feats_pm is None else torch.cat([xyz, feats_pm], dim=2)
            outs = []
            for entry in packed:
                cur = x
                for li, layer in enumerate(entry["layers"]):
                    last = li == len(entry["layers"]) - 1
                    cur = fz.linear(cur, layer, pool=N if last else 1)
                outs.append(cur.view(B, 1, -1))
            return None, (outs[0] if len(outs) == 1 else torch.cat(outs, dim=2))

        if new_xyz is None:
            _, new_xyz = fz.fps_gather(xyz, self.npoint, fps_ordered=fps_ordered)
        M = new_xyz.shape[1]
        if len(self.groupers) == 2:
            g0, g1 = self.groupers
            idxs = fz.ball_query_dual(xyz, new_xyz, g0.radius, g0.nsample, g1.radius, g1.nsample)
        else:
            idxs = [fz.ball_query_single(xyz, new_xyz, g.radius, g.nsample) for g in self.groupers]
        c_total = sum(e["layers"][-1].cout for e in packed)
        out = torch.empty((B, M, c_total), dtype=torch.float32, device=xyz.device)
        out2 = out.view(B * M, c_total)
        col = 0
        for g, idx, entry in zip(self.groupers, idxs, packed):
            layers = entry["layers"]
            c1 = layers[0].cout
            if h_first is not None:
                assert len(packed) == 1 and h_first.shape == (B * N, c1), (h_first.shape, B * N, c1)
                h = h_first
            elif entry["first_f"] is not None:
                h = fz.linear(feats_pm, entry["first_f"])                       # (B*N, c1) per-point half of layer 1
            else:
                h = entry["b1"].unsqueeze(0).expand(B * N, c1).contiguous()
            cout = layers[-1].cout
            dst = out2[:, col:col + cout]
            if len(layers) == 3 and (fz.sa_fused_t_supported(layers[1], layers[2], g.nsample)
                                     or fz.sa_fused_supported(layers[1], layers[2], g.nsample)):
                fz.sa_fused_tc(h, idx, xyz, new_xyz, entry["wxyz"], layers[1], layers[2], dst)
            elif len(layers) == 2:
                fz.sa_group_linear(h, idx, xyz, new_xyz, entry["wxyz"], layers[1], out=dst, pool=g.nsample)
            else:
                cur = fz.sa_group_linear(h, idx, xyz, new_xyz, entry["wxyz"], layers[1])
                for li in range(2, len(layers)):
                    last = li == len(layers) - 1
                    cur = fz.linear(cur, layers[li], out=dst if last else None, pool=g.nsample if last else 1)
            col += cout
        return new_xyz, out

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B, sum C_out, npoint)
        (pointnet2_modules.py:19-55)."""
        if self._can_fuse(xyz):
            feats_pm = features.transpose(1, 2).contiguous() if features is not None else None
            new_xyz, out = self.forward_pm(xyz, feats_pm, new_xyz)
            return new_xyz, out.transpose(1, 2).contiguous()

        new_features_list = []
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if new_xyz is None:
            new_xyz = pointnet2_utils.gather_operation(
                xyz_flipped, pointnet2_utils.furthest_point_sample(xyz, self.npoint)
            ).transpose(1, 2).contiguous() if self.npoint is not None else None
        for i in range(len(self.groupers)):
            new_features = self.groupers[i](xyz, new_xyz, features)  # (B, C, npoint, nsample)
            new_features = self.mlps[i](new_features)
            if self.pool_method == 'max_pool':
                new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            elif self.pool_method == 'avg_pool':
                new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            else:
                raise NotImplementedError
            new_features_list.append(new_features.squeeze(-1))
        return new_xyz, torch.cat(new_features_list, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (pointnet2_modules.py:58-94)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], bn: bool = True,
                 use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for i in range(len(radii)):
            self.groupers.append(
                pointnet2_utils.QueryAndGroup(radii[i], nsamples[i], use_xyz=use_xyz)
                if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            mlp_spec = mlps[i]
            if use_xyz:
                mlp_spec[0] += 3  # in place, like the reference (:88-89): callers see the +3
            self.mlps.append(pt_utils.SharedMLP(mlp_spec, bn=bn, instance_norm=instance_norm))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:95-113)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)


class PointnetFPModule(pt_utils.PackedCacheMixin, nn.Module):
    """Feature propagation (pointnet2_modules.py:116-156)."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)
        self.fused = True
        self._packed = None

    def forward_pm(self, unknown, known, unknow_feats_pm, known_feats_pm):
        """point-major variant: unknow_feats_pm (B,n,C1) or None, known_feats_pm (B,m,C2) -> (B,n,mlp[-1])."""
        if not self._packed_valid():
            self._store_packed(fz.pack_sequential(self.mlp))
        B, n, _ = unknown.shape
        c2 = known_feats_pm.shape[2]
        c1 = unknow_feats_pm.shape[2] if unknow_feats_pm is not None else 0
        # skip connection: when both halves are whole 64-channel K-blocks the first layer reads cat[interpolated, skip] from
        # its two sources (fused.linear_cat) and the copy of the skip features into a concatenation buffer disappears
        two_src = c1 > 0 and c1 % 64 == 0 and c2 % 64 == 0 and fz.MLP_ENGINE == "tc"
        x = torch.empty((B, n, c2 if two_src else c2 + c1), dtype=torch.float32, device=unknown.device)
        x2 = x.view(B * n, -1)
        # three_nn's squared distances go straight into the interpolation kernel, which forms the weights
        # 1 / (sqrt(d2) + 1e-8) / sum like the torch statements of forward() below (bit-identical, one launch for five)
        m = known.shape[1]
        dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
        pointnet2_utils.pointnet2.three_nn_wrapper(B, n, m, unknown, known, dist2, idx)
        fz.three_interpolate_pm_d2(known_feats_pm, idx, dist2, x2[:, :c2])
        if two_src:
            cur = fz.linear_cat(x2, unknow_feats_pm.reshape(B * n, c1), self._packed[0])
            rest = self._packed[1:]
        else:
            if c1:
                x[:, :, c2:] = unknow_feats_pm
            cur, rest = x2, self._packed
        for layer in rest:
            cur = fz.linear(cur, layer)
        return cur.view(B, n, -1)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        if self.fused and not self.training and unknown.is_cuda and known is not None and pt_utils.foldable(self.mlp):
            out = self.forward_pm(unknown, known,
                                  unknow_feats.transpose(1, 2).contiguous() if unknow_feats is not None else None,
                                  known_feats.transpose(1, 2).contiguous())
            return out.transpose(1, 2).contiguous()
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)  # (B, C2 + C1, n)
        else:
            new_features = interpolated_feats
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
