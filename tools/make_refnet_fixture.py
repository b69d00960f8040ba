"""Golden vectors of the REFERENCE network itself for the GPU box: tests/golden/refnet_forward.npz.

tools/refnet_cpu.py runs the reference's unmodified PointRCNN code (lib/net/*.py, pointnet2_lib/pointnet2/*.py,
lib/rpn/proposal_layer.py ...) on the CPU with the CUDA extensions replaced by the C restatements of their kernels.
This script feeds it seeded random-init weights (the product's model has the same state dict, so the same parameters
load into both) and two synthetic 8192-point scenes, and stores what the GPU parity test needs, subsampled to stay
small: rpn_cls (all points), rpn_reg / backbone_features on every 32nd point, seg_result, rois, roi_scores_raw,
rcnn_cls, rcnn_reg.   python tools/make_refnet_fixture.py   (build container only; ~15 s)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
OUT = os.path.join(ROOT, "tests", "golden", "refnet_forward.npz")
B, N, STRIDE, CLOUD_SEED = 2, 8192, 32, 41


def seeded_model(device):
    """default.yaml PointRCNN, torch.manual_seed(0) initialisation, BatchNorm running statistics perturbed from a
    seeded generator (random-init BN is the identity: folding would go untested).  Built on the CPU, then moved."""
    from conftest import load
    load("config").use_default_yaml("rcnn")
    torch.manual_seed(0)
    net = load("net.point_rcnn").PointRCNN(num_classes=2, use_xyz=True, mode="TEST").eval()
    g = torch.Generator(device="cpu").manual_seed(3)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    # focal-loss initialisation puts every point far below the 0.3 foreground threshold; shift the score head so that
    # about half of the points are foreground and the segmentation-mask channel of the RCNN input is exercised
    with torch.no_grad():
        net.rpn.rpn_cls_layer[2].conv.bias.fill_(0.5)
    return net.to(device)


def scenes():
    from conftest import load
    return torch.from_numpy(load("synthetic").make_clouds("lidar", B, N, seed=CLOUD_SEED))


def main():
    import refnet_cpu as rn
    model = seeded_model("cpu")
    ref = rn.build_reference_model(model.state_dict())
    out = rn.reference_forward(ref, scenes())
    keep = {
        "rpn_cls": out["rpn_cls"].numpy(),
        "rpn_reg_sub": out["rpn_reg"][:, ::STRIDE].contiguous().numpy(),
        "backbone_features_sub": out["backbone_features"][:, :, ::STRIDE].contiguous().numpy(),
        "seg_result": out["seg_result"].numpy().astype(np.uint8),
        "rois": out["rois"].numpy(), "roi_scores_raw": out["roi_scores_raw"].numpy(),
        "rcnn_cls": out["rcnn_cls"].numpy(), "rcnn_reg": out["rcnn_reg"].numpy(),
    }
    np.savez_compressed(OUT, **keep)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", int(out["seg_result"].sum()), "foreground points,",
          int((out["rois"].abs().sum(-1) > 0).sum()), "non-empty rois")


if __name__ == "__main__":
    main()
