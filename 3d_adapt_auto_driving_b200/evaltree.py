"""Stage the directory tree in which the reference's UNMODIFIED pointrcnn/tools/eval_rcnn.py runs on
this package.

eval_rcnn.py locates everything relative to itself: `import _init_path` puts ../ , ../lib/datasets
and ../lib/net on sys.path (tools/_init_path.py), it imports lib.net.point_rcnn, lib.datasets.
kitti_rcnn_dataset, tools.train_utils.train_utils, lib.utils.{bbox_transform,kitti_utils},
lib.utils.iou3d.iou3d_utils, lib.config and tensorboardX (eval_rcnn.py:1-23), reads its data from
<tools/..>/multi_data/<dataset> (:854) and is started with cwd = tools/.  make_eval_tree() writes
that skeleton: every module is a two-line shim re-exporting the package module of the same role
(so `lib.config.cfg` IS the package's cfg object), plus tensorboardX / cfg-yaml stand-ins; the
script itself is copied byte-for-byte from a reference checkout the caller points at and checked
against the sha256 recorded in SURVEY.md -- no reference source lives in this repository."""
import hashlib
import os
import shutil

PKG = __name__.rsplit('.', 1)[0]
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EVAL_RCNN_SHA256 = "6485adf662b0621dabb86d8a4d90f058e1e8618243155a4fb9994c219cdb846f"

_SHIMS = {
    "lib/config.py": "config",
    "lib/net/point_rcnn.py": "net.point_rcnn",
    "lib/net/rpn.py": "net.rpn",
    "lib/net/rcnn_net.py": "net.rcnn_net",
    "lib/net/pointnet2_msg.py": "net.pointnet2_msg",
    "lib/rpn/proposal_layer.py": "proposal_layer",
    "lib/datasets/kitti_dataset.py": "datasets.kitti_dataset",
    "lib/datasets/kitti_rcnn_dataset.py": "datasets.kitti_rcnn_dataset",
    "lib/utils/bbox_transform.py": "bbox_transform",
    "lib/utils/kitti_utils.py": "kitti_utils",
    "lib/utils/calibration.py": "calibration",
    "lib/utils/object3d.py": "object3d",
    "lib/utils/iou3d/iou3d_utils.py": "iou3d_utils",
    "lib/utils/roipool3d/roipool3d_utils.py": "roipool3d_utils",
    "tools/train_utils/train_utils.py": "train_utils",
    "pointnet2_lib/pointnet2/pointnet2_utils.py": "pointnet2_utils",
    "pointnet2_lib/pointnet2/pointnet2_modules.py": "pointnet2_modules",
    "pointnet2_lib/pointnet2/pytorch_utils.py": "pytorch_utils",
}

_SHIM_BODY = '''"""shim: the role of this reference module is played by {pkg}.{mod}"""
import importlib as _il
_m = _il.import_module("{pkg}.{mod}")
globals().update({{k: v for k, v in vars(_m).items() if not (k.startswith("__") and k.endswith("__"))}})
'''

# lib/net/point_rcnn.py of the tree: the package's class with the whole-forward CUDA graph switched on (the script drives the
# model from one Python thread: ~170 eager launches per batch cost it more host time than the GPU needs to execute them)
_POINT_RCNN_SHIM = '''"""shim: the role of this reference module is played by {pkg}.net.point_rcnn (graph replay of the inference forward)"""
import importlib as _il
_m = _il.import_module("{pkg}.net.point_rcnn")
globals().update({{k: v for k, v in vars(_m).items() if not (k.startswith("__") and k.endswith("__"))}})


class PointRCNN(_m.PointRCNN):
    graph_forward = __import__("os").environ.get("PN2_MODEL_GRAPH", "1") != "0"
'''

_INIT_PATH = '''import os, sys
sys.path.insert(0, {repo!r})                       # the package
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '../'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '../lib/datasets'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '../lib/net'))
'''

_TENSORBOARDX = '''"""stand-in for tensorboardX (absent offline): eval_rcnn.py only writes scalars in --eval_all mode."""
class SummaryWriter(object):
    def __init__(self, *a, **k): pass
    def add_scalar(self, *a, **k): pass
    def flush(self): pass
    def close(self): pass
'''


def default_yaml_text():
    """tools/cfgs/default.yaml restricted to the keys inference reads (config._DEFAULT_YAML_INFERENCE)."""
    import importlib
    import yaml
    cfgm = importlib.import_module(PKG + ".config")
    return yaml.safe_dump(cfgm._DEFAULT_YAML_INFERENCE, default_flow_style=None)


def make_eval_tree(dest, eval_rcnn_src, check_sha=True):
    """-> <dest>/pointrcnn ; run `python eval_rcnn.py ...` with cwd = <dest>/pointrcnn/tools."""
    with open(eval_rcnn_src, "rb") as f:
        data = f.read()
    if check_sha and hashlib.sha256(data).hexdigest() != EVAL_RCNN_SHA256:
        raise RuntimeError("%s is not the reference's eval_rcnn.py (sha256 mismatch)" % eval_rcnn_src)
    root = os.path.join(dest, "pointrcnn")
    for rel, mod in _SHIMS.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write((_POINT_RCNN_SHIM if rel == "lib/net/point_rcnn.py" else _SHIM_BODY).format(pkg=PKG, mod=mod))
    for d, _, _ in list(os.walk(root)):
        init = os.path.join(d, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    tools = os.path.join(root, "tools")
    os.makedirs(os.path.join(tools, "cfgs"), exist_ok=True)
    with open(os.path.join(tools, "_init_path.py"), "w") as f:
        f.write(_INIT_PATH.format(repo=REPO_ROOT))
    os.makedirs(os.path.join(tools, "tensorboardX"), exist_ok=True)
    with open(os.path.join(tools, "tensorboardX", "__init__.py"), "w") as f:
        f.write(_TENSORBOARDX)
    with open(os.path.join(tools, "cfgs", "default.yaml"), "w") as f:
        f.write(default_yaml_text())
    shutil.copyfile(eval_rcnn_src, os.path.join(tools, "eval_rcnn.py"))
    return root
