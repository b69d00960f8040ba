"""The reference's UNMODIFIED pointrcnn/tools/eval_rcnn.py runs on this package (BASELINE.json north_star,
SURVEY.md 8b): the script is staged byte-for-byte (sha256-checked) in a shim tree (evaltree.py) next
to a synthetic KITTI dataset and a seeded checkpoint, executed as a subprocess with cwd = tools/, and
the KITTI result files it writes are compared with the detections of the package's own batched
Detector on the same sampled clouds."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

pytestmark = pytest.mark.gpu
SCRIPT = os.path.join(ROOT, "oracle", "_ref", "eval_rcnn.py")


def run_script_and_compare(cuda, tmp_path, model, ckpt_path, n_scenes=6, bs=3):
    """Stage the shim tree + a synthetic data set, run the unmodified eval_rcnn.py on `ckpt_path` as a subprocess, and
    compare every KITTI result line with the detections of the package's batched Detector holding `model` (the same
    weights) on the same sampled clouds.  -> number of boxes compared."""
    et, sk, inf = load("evaltree"), load("synthetic_kitti"), load("inference")
    cfgm = load("config")
    root = et.make_eval_tree(str(tmp_path), SCRIPT)
    data_root = sk.make_dataset(root, name="kitti", n_scenes=n_scenes, split="val", seed=666)
    out_dir = tmp_path / "out"
    epoch = os.path.basename(ckpt_path).split("_")[-1].split(".")[0]
    cmd = [sys.executable, "eval_rcnn.py", "--cfg_file", "cfgs/default.yaml", "--eval_mode", "rcnn", "--ckpt",
           str(ckpt_path), "--batch_size", str(bs), "--workers", "0", "--output_dir", str(out_dir)]
    r = subprocess.run(cmd, cwd=os.path.join(root, "tools"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    final = out_dir / "eval" / ("epoch_%s" % epoch) / "val" / "final_result" / "data"
    files = sorted(os.listdir(str(final)))
    assert files == ["%06d.txt" % i for i in range(n_scenes)]

    # the same clouds through the package's batched Detector: eval_one_epoch_joint seeds np.random with 666
    # (eval_rcnn.py:467) and, with --workers 0, the dataset draws from that stream in sample order
    cfgm.use_default_yaml("rcnn")
    ds = load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=data_root, npoints=16384, split="val", mode="EVAL",
                                                              classes="Car", far_points=4000)
    np.random.seed(666)
    det = inf.Detector(model, cuda, use_graph=False)
    ku = load("kitti_utils")
    total = 0
    for b0 in range(0, n_scenes, bs):
        batch = ds.collate_batch([ds[i] for i in range(b0, b0 + bs)])
        rec, cnt = det.detect(torch.from_numpy(batch["pts_input"]).float())
        for k, (boxes, scores) in enumerate(inf.records_to_lists(rec, cnt)):
            sid = int(batch["sample_id"][k])
            lines = [l.split() for l in open(str(final / ("%06d.txt" % sid))).read().splitlines()]
            # save_kitti_format (eval_rcnn.py:76-101) drops boxes wider / taller than 80 % of the image
            calib = ds.get_calib(sid)
            img_boxes, _ = calib.corners3d_to_img_boxes(ku.boxes3d_to_corners3d(boxes)) if len(boxes) else (np.zeros((0, 4)), None)
            shape = ds.get_image_shape(sid)
            x1, y1 = np.clip(img_boxes[:, 0], 0, shape[1] - 1), np.clip(img_boxes[:, 1], 0, shape[0] - 1)
            x2, y2 = np.clip(img_boxes[:, 2], 0, shape[1] - 1), np.clip(img_boxes[:, 3], 0, shape[0] - 1)
            valid = np.logical_and(x2 - x1 < shape[1] * 0.8, y2 - y1 < shape[0] * 0.8)
            assert len(lines) == int(valid.sum()), (sid, len(lines), int(valid.sum()))
            for line, bx, sc in zip(lines, boxes[valid], scores[valid]):
                assert line[0] == "Car"
                got = np.array([float(v) for v in line[8:16]])            # h w l x y z ry score
                want = np.array([bx[3], bx[4], bx[5], bx[0], bx[1], bx[2], bx[6], sc])
                np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
            total += len(lines)
    return total


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="oracle/_ref/eval_rcnn.py not staged (needs /root/reference at build time)")
def test_unmodified_eval_rcnn_runs_and_matches_detector(cuda, tmp_path):
    inf, tu = load("inference"), load("train_utils")
    model = inf.build_model(seed=0, device=cuda)
    # random-init heads score everything below the 0.3 threshold; bias the RCNN score so that boxes survive
    with torch.no_grad():
        model.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
    ckpt_dir = tmp_path / "ckpt"
    ckpt_dir.mkdir()
    tu.save_checkpoint(tu.checkpoint_state(model, None, 1, 1), filename=str(ckpt_dir / "checkpoint_epoch_1"))
    assert run_script_and_compare(cuda, tmp_path, model, ckpt_dir / "checkpoint_epoch_1.pth") > 0


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="oracle/_ref/eval_rcnn.py not staged (needs /root/reference at build time)")
def test_unmodified_eval_rcnn_rpn_mode(cuda, tmp_path):
    """`--eval_mode rpn --test --save_result` (eval_rcnn.py:120-262, eval_one_epoch_rpn): the unmodified script drives the
    RPN-only model of this package -- backbone, heads, proposal layer -- and writes the proposals in KITTI format plus the
    per-point segmentation; the proposals equal those of the package's own model on the same sampled clouds."""
    import glob
    et, sk, inf, tu, cfgm, ku = (load("evaltree"), load("synthetic_kitti"), load("inference"), load("train_utils"), load("config"),
                                 load("kitti_utils"))
    n_scenes, bs = 4, 2
    try:
        model = inf.build_model(seed=0, eval_mode="rpn", device=cuda)
        ckpt_dir = tmp_path / "ckpt"
        ckpt_dir.mkdir()
        tu.save_checkpoint(tu.checkpoint_state(model, None, 1, 1), filename=str(ckpt_dir / "checkpoint_epoch_1"))
        root = et.make_eval_tree(str(tmp_path), SCRIPT)
        data_root = sk.make_dataset(root, name="kitti", n_scenes=n_scenes, split="val", seed=666)
        out_dir = tmp_path / "out"
        cmd = [sys.executable, "eval_rcnn.py", "--cfg_file", "cfgs/default.yaml", "--eval_mode", "rpn", "--test", "--save_result",
               "--ckpt", str(ckpt_dir / "checkpoint_epoch_1.pth"), "--batch_size", str(bs), "--workers", "0",
               "--output_dir", str(out_dir)]
        r = subprocess.run(cmd, cwd=os.path.join(root, "tools"), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
        det_dirs = glob.glob(os.path.join(str(out_dir), "**", "detections", "data"), recursive=True)
        assert len(det_dirs) == 1, det_dirs
        final = det_dirs[0]
        assert sorted(os.listdir(final)) == ["%06d.txt" % i for i in range(n_scenes)]
        seg_dir = os.path.join(os.path.dirname(os.path.dirname(final)), "seg_result")
        assert sorted(os.listdir(seg_dir)) == ["%06d.npy" % i for i in range(n_scenes)]

        # the same clouds through the package's model: eval_one_epoch_rpn seeds np.random with 1024 (eval_rcnn.py:121)
        cfgm.use_default_yaml("rpn")
        ds = load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=data_root, npoints=16384, split="val", mode="TEST",
                                                                  classes="Car", far_points=4000)
        np.random.seed(1024)
        total = 0
        for b0 in range(0, n_scenes, bs):
            batch = ds.collate_batch([ds[i] for i in range(b0, b0 + bs)])
            with torch.no_grad():
                out = model({"pts_input": torch.from_numpy(batch["pts_input"]).float().to(cuda)})
                rois, roi_scores = model.rpn.proposal_layer(out["rpn_cls"][:, :, 0], out["rpn_reg"], out["backbone_xyz"])
                seg = (torch.sigmoid(out["rpn_cls"][:, :, 0]) > cfgm.cfg.RPN.SCORE_THRESH).long()
            for k in range(rois.shape[0]):
                sid = int(batch["sample_id"][k])
                boxes, scores = rois[k].cpu().numpy(), roi_scores[k].cpu().numpy()
                lines = [l.split() for l in open(os.path.join(final, "%06d.txt" % sid)).read().splitlines()]
                calib = ds.get_calib(sid)
                img_boxes, _ = calib.corners3d_to_img_boxes(ku.boxes3d_to_corners3d(boxes))
                shape = ds.get_image_shape(sid)
                x1, y1 = np.clip(img_boxes[:, 0], 0, shape[1] - 1), np.clip(img_boxes[:, 1], 0, shape[0] - 1)
                x2, y2 = np.clip(img_boxes[:, 2], 0, shape[1] - 1), np.clip(img_boxes[:, 3], 0, shape[0] - 1)
                valid = np.logical_and(x2 - x1 < shape[1] * 0.8, y2 - y1 < shape[0] * 0.8)
                assert len(lines) == int(valid.sum()), (sid, len(lines), int(valid.sum()))
                for line, bx, sc in zip(lines, boxes[valid], scores[valid]):
                    got = np.array([float(v) for v in line[8:16]])            # h w l x y z ry score
                    want = np.array([bx[3], bx[4], bx[5], bx[0], bx[1], bx[2], bx[6], sc])
                    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
                saved = np.load(os.path.join(seg_dir, "%06d.npy" % sid))              # (N, 4) float16: xyz | predicted class
                assert np.array_equal(saved[:, 3].astype(np.int64), seg[k].cpu().numpy())
                total += len(lines)
        assert total > 0
    finally:
        cfgm.use_default_yaml("rcnn")
