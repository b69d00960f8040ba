"""GPU data path (csrc/scene_prepare.cu + datasets/gpu_loader.py) against the numpy path of the
dataset mirror (= the reference's get_rpn_sample arithmetic, pointrcnn/lib/datasets/kitti_rcnn_dataset.py:249-342):
bit-exact sampled clouds, and tools/eval_fast.py end to end against the detector fed by the CPU data path."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load, ROOT

pytestmark = pytest.mark.gpu


def _tree(tmp_path, sizes):
    sk = load("synthetic_kitti")
    root = str(tmp_path)
    data_root = os.path.join(root, "multi_data", "kitti")
    train_dir = os.path.join(data_root, "KITTI", "object", "training")
    for sub in ("velodyne", "calib", "label_2", "image_2"):
        os.makedirs(os.path.join(train_dir, sub), exist_ok=True)
    os.makedirs(os.path.join(data_root, "KITTI", "ImageSets"), exist_ok=True)
    rng = np.random.RandomState(7)
    for i, n in enumerate(sizes):
        sk.write_scene(train_dir, i, rng, npoints=n)
    with open(os.path.join(data_root, "KITTI", "ImageSets", "val.txt"), "w") as f:
        f.write("\n".join("%06d" % i for i in range(len(sizes))) + "\n")
    return data_root


def _dataset(data_root):
    load("config").use_default_yaml("rcnn")
    return load("datasets.kitti_rcnn_dataset").KittiRCNNDataset(root_dir=data_root, npoints=16384, split="val", mode="EVAL",
                                                               classes="Car", far_points=4000)


def test_filter_lists_and_counts_match_numpy(cuda, tmp_path):
    """valid points (rect xyz + intensity), their order, near / far lists and counts == the numpy chain."""
    ds = _dataset(_tree(tmp_path, [60000, 5000, 33333]))
    gl = load("datasets.gpu_loader")
    loader = gl.GpuSceneLoader(ds, cuda, batch_size=3)
    np.random.seed(1)
    loader.prepare(range(3))
    buf = next(iter(loader._bufs.values()))
    counts = buf["counts"].cpu().numpy()
    for k in range(3):
        sid = ds.sample_id_list[k]
        calib = ds.get_calib(sid)
        lidar = ds.get_lidar(sid)
        rect = calib.lidar_to_rect(lidar[:, 0:3])
        img, depth = calib.rect_to_img(rect)
        flag = ds.get_valid_flag(rect, img, depth, ds.get_image_shape(sid))
        want = np.concatenate((rect[flag][:, 0:3], lidar[flag][:, 3:4]), axis=1)
        nv = int(counts[k, 0])
        assert nv == len(want) and 0 < nv < len(lidar)
        got = buf["valid"][k, :nv].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))          # bit for bit
        near = np.where(want[:, 2] < 40.0)[0]
        far = np.where(~(want[:, 2] < 40.0))[0]
        assert counts[k, 1] == len(near) and counts[k, 2] == len(far)
        assert np.array_equal(buf["near"][k, :len(near)].cpu().numpy(), near)
        assert np.array_equal(buf["far"][k, :len(far)].cpu().numpy(), far)


@pytest.mark.parametrize("per_scene", [False, True])
def test_batches_equal_collate_of_dataset_items(cuda, tmp_path, per_scene, monkeypatch):
    """same seed -> pts_input / sample_id / gt_boxes3d identical to collate_batch([dataset[i] ...]), in the global
    np.random stream order and with per-scene seeding; scenes with fewer valid points than npoints included."""
    if per_scene:
        monkeypatch.setenv("PN2_PER_SCENE_SEED", "1")
    sizes = [60000, 9000, 40000, 16000, 70000]
    ds = _dataset(_tree(tmp_path, sizes))
    assert ds.per_scene_seed == per_scene
    np.random.seed(666)
    want = [ds.collate_batch([ds[i] for i in idx]) for idx in ([0, 1], [2, 3], [4])]
    gl = load("datasets.gpu_loader")
    np.random.seed(666)
    got = list(gl.GpuSceneLoader(ds, cuda, batch_size=2, with_features=True))
    assert len(got) == 3
    for w, g in zip(want, got):
        assert np.array_equal(g["sample_id"], w["sample_id"])
        assert np.array_equal(g["pts_input"].cpu().numpy().view(np.uint32), w["pts_input"].astype(np.float32).view(np.uint32))
        assert np.array_equal(g["pts_features"].cpu().numpy().view(np.uint32),
                              w["pts_features"].astype(np.float32).view(np.uint32))
        assert np.array_equal(g["gt_boxes3d"], w["gt_boxes3d"])


def test_eval_fast_gpu_and_cpu_data_paths_write_the_same_files(cuda, tmp_path):
    """tools/eval_fast.py: the GPU data path and the reference's numpy data path produce byte-identical KITTI
    result files (same sampled clouds -> same detections), one file per scene of the split."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    eval_fast = importlib.import_module("eval_fast")
    data_root = _tree(tmp_path / "data", [60000, 30000, 9000, 45000, 52000, 38000, 61000])
    inf = load("inference")
    # random-init heads score everything below the 0.3 threshold; bias the RCNN score so that boxes survive
    orig = inf.build_model

    def biased(seed=0, eval_mode="rcnn", device="cuda"):
        m = orig(seed, eval_mode, device)
        with torch.no_grad():
            m.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
        return m
    inf.build_model = biased
    try:
        a = eval_fast.run(data_root, str(tmp_path / "gpu"), batch_size=3, depth=2, gpu_loader=True, log=lambda s: None)
        b = eval_fast.run(data_root, str(tmp_path / "cpu"), batch_size=3, depth=1, gpu_loader=False, log=lambda s: None)
    finally:
        inf.build_model = orig
    fa, fb = sorted(os.listdir(a["final_dir"])), sorted(os.listdir(b["final_dir"]))
    assert fa == fb == ["%06d.txt" % i for i in range(7)]
    assert a["detections"] == b["detections"] > 0
    for f in fa:
        assert open(os.path.join(a["final_dir"], f)).read() == open(os.path.join(b["final_dir"], f)).read(), f
