"""SURVEY 8(f) N3, GPU half: a checkpoint in the reference's on-disk format written by the REFERENCE'S OWN
train_utils.checkpoint_state / save_checkpoint (tools/train_utils/train_utils.py:60-75, imported unmodified from the staged
reference tree) from the REFERENCE'S OWN PointRCNN module, with BatchNorm statistics and affine parameters far from identity,
is evaluated by the unmodified tools/eval_rcnn.py on the package's drop-in tree; the KITTI result files must carry the
detections of the package's Detector holding the same weights (loaded by the package's load_checkpoint), and the reference
network itself (same file, its own kernels) must agree with them box by box."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load
from test_eval_rcnn_dropin_gpu import SCRIPT, run_script_and_compare

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_refnet_fixture as fx             # noqa: E402

pytestmark = pytest.mark.gpu


def _trained_like_(model, seed=3):
    g = torch.Generator().manual_seed(seed)
    n = 0
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            with torch.no_grad():
                m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.2).to(m.running_mean.device))
                m.running_var.copy_((torch.rand(m.running_var.shape, generator=g) * 1.5 + 0.25).to(m.running_var.device))
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) + 0.5).to(m.weight.device))
                m.bias.copy_((torch.randn(m.bias.shape, generator=g) * 0.1).to(m.bias.device))
            n += 1
    return n


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="oracle/_ref/eval_rcnn.py not staged (needs /root/reference at build time)")
def test_reference_written_checkpoint_through_the_unmodified_script(cuda, tmp_path):
    from oracle import refnet_gpu
    if not refnet_gpu.available("legacy"):
        pytest.skip("stock reference tree (baseline/_ref/pointrcnn) or libpn2_legacy.so not available")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        seeded = fx.seeded_model("cpu")                                  # RCNN score head shifted so that boxes survive
        ref = refnet_gpu.Reference(seeded.state_dict(), cuda, backend="legacy")
        assert _trained_like_(ref.model) >= 30
        ckpt_dir = tmp_path / "ckpt"
        ckpt_dir.mkdir()
        with refnet_gpu.reference_imports("legacy"):
            tu_ref = importlib.import_module("train_utils.train_utils")          # the reference's module, not the package's
            assert os.path.realpath(tu_ref.__file__).startswith(os.path.realpath(refnet_gpu.ref_root()))
            tu_ref.save_checkpoint(tu_ref.checkpoint_state(ref.model, None, 42, 999), filename=str(ckpt_dir / "checkpoint_epoch_42"))
        ckpt = ckpt_dir / "checkpoint_epoch_42.pth"
        # the package's loader on the reference's file -> Detector; the unmodified script on the same file -> KITTI files
        model = load("inference").build_model(seed=5, device=cuda)
        it, epoch = load("train_utils").load_checkpoint(model, filename=str(ckpt))
        assert (it, epoch) == (999, 42)
        total = run_script_and_compare(cuda, tmp_path, model, ckpt)
        assert total > 0
        # and the reference network itself on the same weights: its detections for one batch against the Detector's
        import bench
        inf = load("inference")
        pts = fx.scenes()
        want = ref.eval_batch(pts)
        rec, cnt = inf.Detector(model, cuda, use_graph=False, depth=1).detect(pts.pin_memory())
        m = bench.match_detections(want, inf.records_to_lists(rec, cnt))
        print("reference network vs Detector on the reference-written checkpoint:", m)
        assert m["total"] > 0 and m["matched"] >= 0.95 * m["total"] and m["extra"] <= 0.05 * m["total"] + 1
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
