"""Mirror of pointrcnn/pointnet2_lib/pointnet2/pytorch_utils.py: SharedMLP, Conv1d, Conv2d,
BatchNorm1d/2d, FC with the SAME constructor signatures and the SAME child-module names
(`layer{i}` / `conv` / `bn.bn` / `activation`, pytorch_utils.py:22,81-108) so published
checkpoints load by key.  As nn.Modules they run through torch
(training, and the plain fp32 reference the tests compare against); at inference the fused
modules in pointnet2_modules.py read their parameters through `fold_layer` below and run them
on the sm_100a shared-MLP kernels instead.
"""
from typing import List, Tuple

import torch
import torch.nn as nn


class _BNBase(nn.Sequential):
    def __init__(self, in_size, batch_norm=None, name=""):
        super().__init__()
        self.add_module(name + "bn", batch_norm(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class BatchNorm1d(_BNBase):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__(in_size, batch_norm=nn.BatchNorm1d, name=name)


class BatchNorm2d(_BNBase):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, batch_norm=nn.BatchNorm2d, name=name)


class _ConvBase(nn.Sequential):
    """conv(1x1) [+ BN] [+ activation]; bias only when there is no BN (pytorch_utils.py:57)."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=None,
                 batch_norm=None, bias=True, preact=False, name="", instance_norm=False, instance_norm_func=None):
        super().__init__()
        bias = bias and (not bn)
        conv_unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
        init(conv_unit.weight)
        if bias:
            nn.init.constant_(conv_unit.bias, 0)
        norm_size = in_size if preact else out_size
        bn_unit = batch_norm(norm_size) if bn else None
        in_unit = instance_norm_func(norm_size, affine=False, track_running_stats=False) if instance_norm else None

        def add_norm_act():
            if bn:
                self.add_module(name + "bn", bn_unit)
            if activation is not None:
                self.add_module(name + "activation", activation)
            if not bn and instance_norm:
                self.add_module(name + "in", in_unit)

        if preact:
            add_norm_act()
        self.add_module(name + "conv", conv_unit)
        if not preact:
            add_norm_act()


class Conv1d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv1d,
                         batch_norm=BatchNorm1d, bias=bias, preact=preact, name=name, instance_norm=instance_norm,
                         instance_norm_func=nn.InstanceNorm1d)


class Conv2d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int] = (1, 1),
                 stride: Tuple[int, int] = (1, 1), padding: Tuple[int, int] = (0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv2d,
                         batch_norm=BatchNorm2d, bias=bias, preact=preact, name=name, instance_norm=instance_norm,
                         instance_norm_func=nn.InstanceNorm2d)


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = (not first) or (not preact) or (i != 0)
            self.add_module(name + "layer{}".format(i),
                            Conv2d(args[i], args[i + 1], bn=plain and bn, activation=activation if plain else None,
                                   preact=preact, instance_norm=instance_norm))


class FC(nn.Sequential):
    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False, init=None,
                 preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)

        def add_norm_act(size):
            if bn:
                self.add_module(name + "bn", BatchNorm1d(size))
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            add_norm_act(in_size)
        self.add_module(name + "fc", fc)
        if not preact:
            add_norm_act(out_size)


class PackedCacheMixin:
    """Folded (conv + eval BatchNorm) and tensor-core-packed weights are DERIVED data kept in `self._packed`.  They are
    dropped on a train()/eval() mode switch, on load_state_dict, on any Module._apply (.to / .cuda / .half / .float move
    or retype the sources) and whenever a source parameter / buffer was replaced or written in place since packing
    (optimizer step, `p.copy_()` under no_grad, BatchNorm statistics update: torch bumps `_version`).  Not detectable:
    writes through `.data` (they bypass the version counter) -- call invalidate_packed() after those.  A captured CUDA
    graph (inference.Detector) bakes the packed buffers in: re-create the Detector after changing weights."""

    def invalidate_packed(self):
        self._packed = None
        self._packed_key = None

    def _source_modules(self):
        """modules whose parameters / buffers the cache is derived from (default: the whole module)."""
        return (self,)

    def _source_key(self):
        key = []
        for m in self._source_modules():
            key.extend((id(t), t._version, t.device) for t in m.parameters())
            key.extend((id(t), t._version, t.device) for t in m.buffers())
        return tuple(key)

    def _packed_valid(self):
        """True when `self._packed` exists and still describes the current parameters (call before using it)."""
        if getattr(self, "_packed", None) is None:
            return False
        if getattr(self, "_packed_key", None) != self._source_key():
            self.invalidate_packed()
            return False
        return True

    def _store_packed(self, packed):
        self._packed = packed
        self._packed_key = self._source_key()
        return packed

    def train(self, mode=True):
        if bool(mode) != self.training:   # eval() on an eval-mode module (point_rcnn.py:34 does it every forward) keeps the cache
            self.invalidate_packed()
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)


# ---------------------------------------------------------------------------------------------
# inference-time view of a conv block: (W (cout, cin) f32, b (cout) f32, relu: bool) with the
# eval-mode BatchNorm folded in:  y = gamma * (Wx + b0 - mean) / sqrt(var + eps) + beta
# ---------------------------------------------------------------------------------------------
def fold_layer(block):
    """block: a _ConvBase (conv [+ bn.bn] [+ activation], post-activation order only)."""
    names = [n for n, _ in block.named_children()]
    conv = getattr(block, [n for n in names if n.endswith("conv")][0])
    if names.index([n for n in names if n.endswith("conv")][0]) != 0:
        raise NotImplementedError("pre-activation blocks are not on the PointRCNN path")
    if any(n.endswith("in") for n in names):
        raise NotImplementedError("instance norm is not on the PointRCNN path")
    w = conv.weight.detach().reshape(conv.weight.shape[0], -1).to(torch.float32)
    b = conv.bias.detach().to(torch.float32) if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    bn_names = [n for n in names if n.endswith("bn")]
    if bn_names:
        bn = getattr(block, bn_names[0])[0]
        scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach()) * scale + bn.bias.detach()
    acts = [getattr(block, n) for n in names if n.endswith("activation")]
    if any(not isinstance(a, nn.ReLU) for a in acts):
        raise NotImplementedError("only ReLU (or no activation) folds into the fused kernels")
    return w.contiguous(), b.contiguous(), bool(acts)


def foldable(seq):
    """every conv block of an nn.Sequential (SharedMLP / head) is post-activation conv [+ BN] [+ ReLU]: what the fused
    kernels implement; anything else takes the reference-structured path."""
    for m in seq.children():
        if isinstance(m, nn.Dropout):
            continue
        names = [n for n, _ in m.named_children()]
        convs = [n for n in names if n.endswith("conv")]
        if len(convs) != 1 or names.index(convs[0]) != 0 or any(n.endswith("in") for n in names):
            return False
        if any(not isinstance(getattr(m, n), nn.ReLU) for n in names if n.endswith("activation")):
            return False
    return True
