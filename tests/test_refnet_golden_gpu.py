"""The sm_100a path against golden vectors of the REFERENCE network itself (tests/golden/refnet_forward.npz, written
by tools/make_refnet_fixture.py: the reference's unmodified PointRCNN code run on the CPU with its CUDA extensions
replaced by the C restatements of their kernels).  Same seeded weights, same scenes.
  * RPN stage: scores, regression and point features of the fused kernels within 1e-4 of the tensor scale.
  * RCNN stage, teacher-forced with the reference's ROIs and segmentation mask on top of this path's own RPN
    features: classification and regression within 1e-4 of the tensor scale as well.
Measured on a B200: 5.4e-5 / 3.1e-5 / 2.4e-5 (rpn_cls / rpn_reg / backbone_features), 2.4e-5 / 1.4e-5 (rcnn_cls / rcnn_reg).
The proposal layer in between is discrete (top-k, NMS); it is pinned separately against the reference recipe
(test_mlp_modules_gpu.py::test_proposal_layer_vs_reference_nms_composition) and, bit for bit on the CPU, against the
reference's ProposalLayer code (test_refnet_vs_port_cpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_refnet_fixture as fx             # noqa: E402

pytestmark = pytest.mark.gpu


def _close(got, want, rel, what):
    want = torch.as_tensor(want).to(got.device)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    print("%s: max err / scale = %.2e" % (what, err / scale))
    assert err <= rel * scale, "%s: max err %.3e at scale %.3e" % (what, err, scale)


def test_fused_path_matches_reference_network_golden(cuda):
    z = np.load(os.path.join(ROOT, "tests", "golden", "refnet_forward.npz"))
    model = fx.seeded_model(cuda)
    pts = fx.scenes().to(cuda)
    with torch.no_grad():
        rpn = model.rpn({"pts_input": pts})
        assert torch.equal(rpn["backbone_xyz"], pts)
        _close(rpn["rpn_cls"], z["rpn_cls"], 1e-4, "rpn_cls")
        _close(rpn["rpn_reg"][:, ::fx.STRIDE], z["rpn_reg_sub"], 1e-4, "rpn_reg")
        _close(rpn["backbone_features"][:, :, ::fx.STRIDE], z["backbone_features_sub"], 1e-4, "backbone_features")
        # the foreground mask is a threshold on the scores: identical except where a score sits within the error
        # bound of the threshold
        scores = rpn["rpn_cls"][:, :, 0]
        seg = (torch.sigmoid(scores) > 0.3).float()
        gold_seg = torch.from_numpy(z["seg_result"].astype(np.float32)).to(cuda)
        flips = seg != gold_seg
        thr = float(np.log(0.3 / 0.7))
        assert bool(((scores - thr).abs()[flips] < 1e-4 * float(np.abs(z["rpn_cls"]).max())).all())
        assert int(flips.sum()) <= 4
        out = model.rcnn_net({"rpn_xyz": rpn["backbone_xyz"], "rpn_features": rpn["backbone_features"].permute(0, 2, 1),
                              "seg_mask": gold_seg, "roi_boxes3d": torch.from_numpy(z["rois"]).to(cuda),
                              "pts_depth": torch.norm(rpn["backbone_xyz"], p=2, dim=2)})
        _close(out["rcnn_cls"], z["rcnn_cls"], 1e-4, "rcnn_cls")
        _close(out["rcnn_reg"], z["rcnn_reg"], 1e-4, "rcnn_reg")
