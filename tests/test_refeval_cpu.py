"""The REFERENCE pipeline end to end -- its unmodified tools/eval_rcnn.py, KittiRCNNDataset, PointRCNN, post-processing
and save_kitti_format, run on the CPU by tools/refnet_cpu.py -- against the CPU port and the host mirrors of this
repository (dataset mirror -> oracle/cpu_forward.py -> kitti_output.save_kitti_format) on the same synthetic KITTI tree
and checkpoint: the result files must be BYTE-IDENTICAL, and identical to the committed golden files
(tests/golden/refeval/, which the GPU box checks the sm_100a Detector against).  Build container only (~40 s)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import refnet_cpu as rn                      # noqa: E402
import make_refeval_fixture as fx            # noqa: E402

pytestmark = pytest.mark.skipif(not rn.available(), reason="reference tree not present")


def test_reference_pipeline_result_files_equal_port_and_golden(tmp_path):
    from oracle import cpu_forward as cf
    final = fx.run_reference(str(tmp_path))
    names = sorted(os.listdir(final))
    assert names == ["%06d.txt" % i for i in range(fx.N_SCENES)]
    cfgm, ko = load("config"), load("kitti_output")
    cfgm.use_default_yaml("rcnn")
    model = fx.seeded_model("cpu")
    data_root = os.path.join(str(tmp_path), "pointrcnn", "multi_data", "kitti")
    ds = load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=data_root, npoints=16384, split="val", mode="EVAL",
                                                              classes="Car", far_points=4000)
    np.random.seed(666)                                   # eval_one_epoch_joint seeds the sampling stream (eval_rcnn.py:467)
    batch = ds.collate_batch([ds[i] for i in range(fx.N_SCENES)])
    pkg = {"cfg": cfgm.cfg, "decode_bbox_target": load("bbox_transform").decode_bbox_target}
    out = cf.pointrcnn_forward(pkg, model, torch.from_numpy(batch["pts_input"]).float())
    mine = tmp_path / "port"
    mine.mkdir()
    total = 0
    for k, (boxes, scores) in enumerate(cf.postprocess(pkg, out, fx.N_SCENES)):
        sid = int(batch["sample_id"][k])
        ko.save_kitti_format(sid, ds.get_calib(sid), boxes, str(mine), scores, ds.get_image_shape(sid))
        want = open(os.path.join(final, "%06d.txt" % sid)).read()
        assert open(str(mine / ("%06d.txt" % sid))).read() == want, sid
        assert open(os.path.join(fx.GOLD, "%06d.txt" % sid)).read() == want, "golden files are stale: rerun tools/make_refeval_fixture.py"
        total += len(want.splitlines())
    assert total > 100
