"""datasets/kitti_rcnn_dataset.py (the mirror eval_rcnn.py iterates) against the REFERENCE class itself
(pointrcnn/lib/datasets/kitti_rcnn_dataset.py imported from /root/reference; `easydict`, the `roipool3d_cuda`
extension and PyYAML's old load() signature are stubbed -- none of them is on the sample path) on a synthetic KITTI
tree: same samples and same collated batches from the same np.random seed, dtype for dtype.  Build container only.
The mirror leaves out `rpn_cls_label` / `rpn_reg_label` (training labels the evaluation loop never reads)."""
import logging
import os
import sys
import types

import numpy as np
import pytest

from conftest import load

REF = "/root/reference/pointrcnn"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def reference_dataset_module():
    import yaml
    cfgm = load("config")
    before = set(sys.modules)
    ed = types.ModuleType("easydict")
    ed.EasyDict = cfgm.AttrDict
    sys.modules["easydict"] = ed
    sys.modules["roipool3d_cuda"] = types.ModuleType("roipool3d_cuda")
    old_load = yaml.load
    yaml.load = lambda f, Loader=yaml.FullLoader: old_load(f, Loader=Loader)
    sys.path.insert(0, REF)
    try:
        import lib.config as rcfg
        import lib.datasets.kitti_rcnn_dataset as rds
        rcfg.cfg_from_file(os.path.join(REF, "tools", "cfgs", "default.yaml"))
        rcfg.cfg.TAG = "default"
        rcfg.cfg.RPN.ENABLED = rcfg.cfg.RCNN.ENABLED = True          # what eval_rcnn.py --eval_mode rcnn sets
    finally:
        sys.path.remove(REF)
        yaml.load = old_load
        for k in set(sys.modules) - before:                          # the reference's `lib` must not shadow anything later
            if k == "lib" or k.startswith("lib.") or k in ("easydict", "roipool3d_cuda"):
                del sys.modules[k]
    return rds


@pytest.mark.parametrize("mode", ["EVAL", "TEST"])
def test_samples_and_batches_equal(reference_dataset_module, tmp_path, mode):
    cfgm, sk, mine = load("config"), load("synthetic_kitti"), load("datasets.kitti_rcnn_dataset")
    cfgm.use_default_yaml("rcnn")
    root = sk.make_dataset(str(tmp_path), name="kitti", n_scenes=4, split="val", seed=1, npoints=30000)
    kw = dict(npoints=16384, split="val", mode=mode, random_select=True, classes="Car")
    a = reference_dataset_module.KittiRCNNDataset(root, logger=logging.getLogger("ref_ds"), **kw)
    b = mine.KittiRCNNDataset(root, **kw)
    assert len(a) == len(b) == 4
    for seed in (666, 3):
        np.random.seed(seed)
        sa = [a[i] for i in range(len(a))]
        state_a = np.random.get_state()[1].copy()
        np.random.seed(seed)
        sb = [b[i] for i in range(len(b))]
        assert np.array_equal(state_a, np.random.get_state()[1])                  # same number of draws
        for x, y in zip(sa, sb):
            assert set(y) <= set(x) and set(x) - set(y) <= {"rpn_cls_label", "rpn_reg_label"}
            for k in y:
                xa, ya = np.asarray(x[k]), np.asarray(y[k])
                assert xa.dtype == ya.dtype and np.array_equal(xa, ya), (mode, k)
        ca, cb = a.collate_batch(sa), b.collate_batch(sb)
        for k in cb:
            xa, ya = np.asarray(ca[k]), np.asarray(cb[k])
            assert xa.dtype == ya.dtype and np.array_equal(xa, ya), (mode, k)
