"""The reference's OWN network code (lib/net/*.py, pointnet2_lib/pointnet2/*.py, lib/rpn/proposal_layer.py, imported
unmodified from /root/reference and run on the CPU by tools/refnet_cpu.py with the CUDA extensions replaced by the C
restatements of their kernels) against oracle/cpu_forward.py, the CPU port that the GPU parity tests, smoke() and
the bench's cpu_baseline lean on: same weights (the product's state dict loads into the reference model with
strict=True), same scenes -> every output tensor of PointRCNN.forward EQUAL, bit for bit.  Build container only; the
GPU box checks the sm_100a path against golden vectors of the same reference run (tests/test_refnet_golden_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import refnet_cpu as rn                      # noqa: E402
import make_refnet_fixture as fx             # noqa: E402

pytestmark = pytest.mark.skipif(not rn.available(), reason="reference tree not present")

KEYS = ("rpn_cls", "rpn_reg", "backbone_xyz", "backbone_features", "seg_result", "rois", "roi_scores_raw", "rcnn_cls",
        "rcnn_reg")


def test_reference_network_equals_cpu_port_and_golden():
    from oracle import cpu_forward as cf
    model = fx.seeded_model("cpu")
    ref = rn.build_reference_model(model.state_dict())             # strict=True: same keys, same shapes
    pts = fx.scenes()
    want = rn.reference_forward(ref, pts)
    pkg = {"cfg": load("config").cfg, "decode_bbox_target": load("bbox_transform").decode_bbox_target}
    got = cf.pointrcnn_forward(pkg, model, pts)
    assert set(KEYS) <= set(want) and set(KEYS) <= set(got)
    for k in KEYS:
        assert want[k].shape == got[k].shape and torch.equal(want[k], got[k]), k
    assert 0.2 < float(want["seg_result"].mean()) < 0.8            # the mask channel is not trivially constant
    assert int((want["rois"].abs().sum(-1) > 0).sum()) == want["rois"].shape[0] * want["rois"].shape[1]
    # the committed golden file is this very run (regenerate with tools/make_refnet_fixture.py if the recipe changes)
    z = np.load(os.path.join(ROOT, "tests", "golden", "refnet_forward.npz"))
    assert np.array_equal(z["rois"], want["rois"].numpy())
    assert np.allclose(z["rcnn_reg"], want["rcnn_reg"].numpy(), rtol=0, atol=1e-5 * float(np.abs(z["rcnn_reg"]).max()))
