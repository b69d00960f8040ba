"""SURVEY 8(f) N3, CPU half: a checkpoint in the reference's on-disk format, WRITTEN BY THE REFERENCE'S OWN CODE
(tools/train_utils/train_utils.py:60-75 checkpoint_state + save_checkpoint on the reference's own PointRCNN module, imported
unmodified from /root/reference through tools/refnet_cpu.py), with BatchNorm statistics and affine parameters far from
identity -- what a trained model has and a random-init one does not -- loads into the package's model through the package's
load_checkpoint, and the folded weights the sm_100a kernels consume reproduce conv + BN of the reference module.
The GPU half (tests/test_checkpoint_roundtrip_gpu.py) pushes such a file through the unmodified eval_rcnn.py."""
import importlib
import os
import sys

import pytest
import torch

from conftest import ROOT, load

sys.path.insert(0, os.path.join(ROOT, "tools"))
import refnet_cpu as rn                      # noqa: E402
import make_refnet_fixture as fx             # noqa: E402

pytestmark = pytest.mark.skipif(not rn.available(), reason="reference tree not present")


def trained_like_(model, seed=3):
    """in place: BatchNorm running statistics and affine parameters as after training (not mean 0 / var 1 / gamma 1 / beta 0)"""
    g = torch.Generator().manual_seed(seed)
    n = 0
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 1.5 + 0.25)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.num_batches_tracked.fill_(1234)
            n += 1
    return n


def reference_checkpoint(path_without_ext, seed=3):
    """reference model <- the package's seeded weights, perturbed BN, saved by the reference's own train_utils"""
    ref = rn.build_reference_model(fx.seeded_model("cpu").state_dict())
    assert trained_like_(ref, seed) >= 30
    with rn.reference_imports():
        sys.path.insert(0, os.path.join(rn.REF, "tools"))
        try:
            tu = importlib.import_module("train_utils.train_utils")
            tu.save_checkpoint(tu.checkpoint_state(ref, None, 77, 4321), filename=path_without_ext)
        finally:
            sys.path.remove(os.path.join(rn.REF, "tools"))
            for k in [k for k in sys.modules if k == "train_utils" or k.startswith("train_utils.")]:
                del sys.modules[k]
    return ref


def test_reference_written_checkpoint_loads_and_folds(tmp_path):
    ref = reference_checkpoint(str(tmp_path / "checkpoint_epoch_77"))
    ckpt = str(tmp_path / "checkpoint_epoch_77.pth")
    blob = torch.load(ckpt, map_location="cpu")
    assert set(blob) == {"epoch", "it", "model_state", "optimizer_state"} and blob["epoch"] == 77 and blob["it"] == 4321
    model = load("inference").build_model(seed=1, device="cpu")          # different weights before loading
    it, epoch = load("train_utils").load_checkpoint(model, filename=ckpt)
    assert (it, epoch) == (4321, 77)
    want = ref.state_dict()
    got = model.state_dict()
    assert list(got) == list(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # the folded layer the kernels run == the reference's conv + BatchNorm (eval) on the same input, for a BN layer whose
    # statistics are now far from identity
    pu, fz = load("pytorch_utils"), load("fused")
    block = model.rpn.backbone_net.SA_modules[1].mlps[0][1]          # the RPN backbone has BatchNorm (cfg.RPN.USE_BN)
    ref_block = ref.rpn.backbone_net.SA_modules[1].mlps[0][1]
    w, b, relu = pu.fold_layer(block)
    x = torch.randn((1, w.shape[1], 50, 3), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y_ref = ref_block(x)
    y = torch.einsum("oc,bcnk->bonk", w.double(), x.double()) + b.double().view(1, -1, 1, 1)
    y = y.clamp_min(0) if relu else y
    assert float((y.float() - y_ref).abs().max()) <= 2e-5 * float(y_ref.abs().max())
    bns = [m for m in block.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert bns and float(bns[0].running_mean.abs().max()) > 0.1
