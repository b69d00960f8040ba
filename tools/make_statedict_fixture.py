"""Generate tests/golden/pointrcnn_state_dict.json from the REFERENCE model definition.

Runs in the build container (CPU): imports /root/reference/pointrcnn/lib/net/point_rcnn.py with
stub modules for the things that are absent or need a GPU (easydict, the three CUDA
extensions, shapely-free loss utils), applies tools/cfgs/default.yaml + eval_mode 'rcnn' and
records every state-dict key with its shape.  The product's PointRCNN must expose exactly the
same keys and shapes (tests/test_state_dict_compat.py) so the published checkpoints load.
"""
import json
import os
import sys
import types

import torch
import yaml

REF = "/root/reference/pointrcnn"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "pointrcnn_state_dict.json")


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def main():
    ed = types.ModuleType("easydict"); ed.EasyDict = EasyDict; sys.modules["easydict"] = ed
    for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda"):
        sys.modules[name] = types.ModuleType(name)
    sh = types.ModuleType("shapely"); shg = types.ModuleType("shapely.geometry"); shg.Polygon = object
    sys.modules["shapely"] = sh; sys.modules["shapely.geometry"] = shg
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "lib", "net"))
    torch.Tensor.cuda = lambda self, *a, **k: self  # ProposalLayer.__init__ calls .cuda()
    _load = yaml.load
    yaml.load = lambda f, *a, **k: _load(f, Loader=yaml.SafeLoader)  # config.py:188 predates PyYAML 6
    from lib.config import cfg, cfg_from_file
    cfg_from_file(os.path.join(REF, "tools", "cfgs", "default.yaml"))
    cfg.RCNN.ENABLED = True
    cfg.RPN.ENABLED = cfg.RPN.FIXED = True
    from lib.net.point_rcnn import PointRCNN
    model = PointRCNN(num_classes=2, use_xyz=True, mode="TEST")
    sd = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(OUT, "w") as f:
        json.dump(sd, f, indent=0, sort_keys=True)
    print(len(sd), "keys ->", OUT)


if __name__ == "__main__":
    main()
