"""The sm_100a Detector against result files written by the REFERENCE pipeline itself (tests/golden/refeval/, produced
by tools/make_refeval_fixture.py: the reference's unmodified eval_rcnn.py + dataset + network + writer on the CPU with
the C restatements of its kernels).  Same synthetic KITTI tree (regenerated from its seed), same seeded weights, same
np.random sampling stream.
The GPU features differ from the CPU run by <= 5.4e-5 of scale (test_refnet_golden_gpu.py), so a box whose score,
overlap or rank sits within that distance of a threshold may legitimately flip; everything else must agree.  Bar: per
scene the box count differs by at most 2 and at least 95 % of the reference's boxes have a Detector box with the same
geometry and score within 2e-3 (the files hold 4 decimals).
Calibration of the bar: adding uniform noise of 3e-5 of the tensor scale to the RCNN head outputs of the CPU run itself
moves 0-4 (mean 1.6) of the 140 boxes by more than 2e-3 over 40 trials -- bin-based decoding takes an argmax over bin
logits, and random-init heads have near-tied bins -- so the B200 result (139 / 140) is what the feature tolerance
predicts, not a logic difference."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_refeval_fixture as fx            # noqa: E402

pytestmark = pytest.mark.gpu


def test_detector_matches_reference_pipeline_result_files(cuda, tmp_path):
    inf, cfgm, ko = load("inference"), load("config"), load("kitti_output")
    cfgm.use_default_yaml("rcnn")
    data_root = fx.make_dataset(str(tmp_path))
    model = fx.seeded_model(cuda)
    ds = load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=data_root, npoints=16384, split="val", mode="EVAL",
                                                              classes="Car", far_points=4000)
    np.random.seed(666)
    batch = ds.collate_batch([ds[i] for i in range(fx.N_SCENES)])
    det = inf.Detector(model, cuda, use_graph=False)
    rec, cnt = det.detect(torch.from_numpy(batch["pts_input"]).float())
    matched = total = 0
    for k, (boxes, scores) in enumerate(inf.records_to_lists(rec, cnt)):
        sid = int(batch["sample_id"][k])
        lines = ko.kitti_lines(ds.get_calib(sid), boxes, scores, ds.get_image_shape(sid))
        got = np.array([[float(v) for v in l.split()[3:16]] for l in lines]).reshape(-1, 13)
        gold = [l.split() for l in open(os.path.join(fx.GOLD, "%06d.txt" % sid)).read().splitlines()]
        want = np.array([[float(v) for v in l[3:16]] for l in gold]).reshape(-1, 13)
        assert all(l[0] == "Car" for l in gold)
        assert abs(len(got) - len(want)) <= 2, (sid, len(got), len(want))
        for w in want:
            # columns: alpha, image box (4), h w l, x y z, ry, score; the image box amplifies by the focal length
            err = np.abs(got - w)
            err[:, 1:5] /= 100.0
            matched += int(bool((err.max(axis=1) <= 2e-3).any())) if len(got) else 0
        total += len(want)
    print("matched %d / %d reference boxes" % (matched, total))
    assert total > 100 and matched >= 0.95 * total
