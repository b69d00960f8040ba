"""The drop-in tree for the unmodified eval_rcnn.py (evaltree.py) without a GPU: the staged script is the
reference's byte for byte (sha256), a wrong file is refused, every shim module resolves to the package module of the
same role when imported the way eval_rcnn.py does (cwd = tools/, `import _init_path`), and the dataset mirror serves
eval_rcnn.py's EVAL-mode calls on a synthetic KITTI tree."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import load, ROOT

SCRIPT = os.path.join(ROOT, "oracle", "_ref", "eval_rcnn.py")
pytestmark = pytest.mark.skipif(not os.path.exists(SCRIPT), reason="oracle/_ref/eval_rcnn.py not staged (needs /root/reference at build time)")


def test_tree_layout_sha_and_shims(tmp_path):
    et, sk = load("evaltree"), load("synthetic_kitti")
    root = et.make_eval_tree(str(tmp_path), SCRIPT)
    staged = os.path.join(root, "tools", "eval_rcnn.py")
    assert hashlib.sha256(open(staged, "rb").read()).hexdigest() == et.EVAL_RCNN_SHA256
    bad = tmp_path / "not_eval_rcnn.py"
    bad.write_text("print('hello')\n")
    with pytest.raises(RuntimeError):
        et.make_eval_tree(str(tmp_path / "x"), str(bad))
    for rel in ("lib/net/point_rcnn.py", "lib/datasets/kitti_rcnn_dataset.py", "lib/utils/iou3d/iou3d_utils.py",
                "pointnet2_lib/pointnet2/pointnet2_utils.py", "tools/train_utils/train_utils.py", "tools/cfgs/default.yaml",
                "tools/_init_path.py", "tools/tensorboardX/__init__.py"):
        assert os.path.exists(os.path.join(root, rel)), rel
    data_root = sk.make_dataset(root, name="kitti", n_scenes=3, split="val", seed=1, npoints=20000)
    # import the modules exactly like eval_rcnn.py:1-23 does, in a fresh interpreter with cwd = tools/
    code = r'''
import _init_path, sys, numpy as np
from lib.config import cfg, cfg_from_file
from lib.datasets.kitti_rcnn_dataset import KittiRCNNDataset
import lib.utils.kitti_utils as kitti_utils
from lib.utils.bbox_transform import decode_bbox_target
import tools.train_utils.train_utils as train_utils
from tensorboardX import SummaryWriter
import lib.utils.iou3d.iou3d_utils as iou3d_utils
cfg_from_file("cfgs/default.yaml")
cfg.RCNN.ENABLED = True; cfg.RPN.ENABLED = cfg.RPN.FIXED = True
ds = KittiRCNNDataset(root_dir=sys.argv[1], npoints=cfg.RPN.NUM_POINTS, split=cfg.TEST.SPLIT, mode="EVAL", random_select=True,
                      classes=cfg.CLASSES, far_points=4000)
np.random.seed(666)
b = ds.collate_batch([ds[0], ds[1]])
assert b["pts_input"].shape == (2, 16384, 3) and b["pts_input"].dtype == np.float32
assert b["gt_boxes3d"].shape[0] == 2 and b["gt_boxes3d"].shape[2] == 7
assert list(b["sample_id"]) == [0, 1] and ds.num_class == 2 and len(ds) == 3
calib = ds.get_calib(0); shape = ds.get_image_shape(0)
boxes, corners = calib.corners3d_to_img_boxes(kitti_utils.boxes3d_to_corners3d(b["gt_boxes3d"][0]))
assert boxes.shape[1] == 4 and shape == (375, 1242, 3)
assert KittiRCNNDataset.__module__.endswith("datasets.kitti_rcnn_dataset") and callable(decode_bbox_target)
print("ok")
'''
    r = subprocess.run([sys.executable, "-c", code, data_root], cwd=os.path.join(root, "tools"), capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
