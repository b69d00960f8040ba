"""The native data path of datasets/kitti_rcnn_dataset.py (pn2_scene_filter_host_f32 + the MT19937 replay, what the
DataLoader workers of the unmodified eval_rcnn.py run) against the numpy path of the same class, which
tests/test_dataset_vs_reference_cpu.py pins to the reference class: identical samples, dtype for dtype, and np.random left in
the identical state -- for sweeps with many invisible points, scenes smaller than the point budget (padding branch), far
bands above and below the cap, points exactly on the camera plane, the global stream and per-scene seeds."""
import os

import numpy as np
import pytest

from conftest import load


def _tree(tmp_path, sizes):
    sk = load("synthetic_kitti")
    root = sk.make_dataset(str(tmp_path), name="kitti", n_scenes=len(sizes), split="val", seed=3, npoints=1000)
    velo = os.path.join(root, "KITTI", "object", "training", "velodyne")
    rng = np.random.RandomState(11)
    for i, (visible, invisible, far_shift) in enumerate(sizes):
        pts = load("synthetic").lidar_cloud(rng, visible)
        pts[:, 2] += far_shift                                     # push part of the scene beyond 40 m
        if invisible:
            back = load("synthetic").uniform_cloud(rng, invisible)
            back[:, 2] = -back[:, 2] - 1.0
            pts = np.concatenate([pts, back])[rng.permutation(visible + invisible)]
        raw = np.concatenate([sk._rect_to_velo(pts), rng.random_sample((len(pts), 1))], axis=1).astype(np.float32)
        raw[0, :3] = 0.0                                            # a return at the sensor origin
        raw.tofile(os.path.join(velo, "%06d.bin" % i))
    return root


@pytest.mark.parametrize("per_scene_seed", [False, True])
def test_native_path_equals_numpy_path(tmp_path, monkeypatch, per_scene_seed):
    cfgm, mod = load("config"), load("datasets.kitti_rcnn_dataset")
    cfgm.use_default_yaml("rcnn")
    # (visible, invisible, z shift): big sweep; fewer points than the budget; tiny scene (more padding than points); a far
    # band above the 4000 cap; a scene whose near band is the small one
    root = _tree(tmp_path, [(24000, 90000, 0.0), (9000, 20000, 0.0), (3000, 0, 0.0), (40000, 1000, 25.0), (30000, 0, 20.0)])
    monkeypatch.setenv("PN2_PER_SCENE_SEED", "1" if per_scene_seed else "0")
    for mode in ("EVAL", "TEST"):
        ds = mod.KittiRCNNDataset(root, npoints=16384, split="val", mode=mode, random_select=True, classes="Car")
        assert ds.per_scene_seed == per_scene_seed
        out = {}
        for native in (False, True):
            monkeypatch.setattr(mod, "NATIVE_DATAPATH", native)
            np.random.seed(666)
            samples = [ds[i] for i in range(len(ds))]
            out[native] = (samples, np.random.get_state())
        (sa, st_a), (sb, st_b) = out[False], out[True]
        assert st_a[0] == st_b[0] and np.array_equal(st_a[1], st_b[1]) and st_a[2:] == st_b[2:]
        for i, (x, y) in enumerate(zip(sa, sb)):
            assert list(x) == list(y)
            for k in x:
                xa, ya = np.asarray(x[k]), np.asarray(y[k])
                assert xa.dtype == ya.dtype and xa.shape == ya.shape and np.array_equal(xa, ya), (mode, i, k)
        assert sa[0]["pts_input"].shape == (16384, 3)
