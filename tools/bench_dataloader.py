"""Throughput of the data path the unmodified eval_rcnn.py uses (torch DataLoader over KittiRCNNDataset, numpy batches), without
any model: scenes/s for 0 / 4 / 8 / 16 workers, native vs numpy per-scene pipeline, shared-memory vs pickled batches."""
import importlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def main():
    import numpy as np
    import torch
    from torch.utils.data import DataLoader
    sk = importlib.import_module(PKG + ".synthetic_kitti")
    cfgm = importlib.import_module(PKG + ".config")
    mod = importlib.import_module(PKG + ".datasets.kitti_rcnn_dataset")
    cfgm.use_default_yaml("rcnn")
    os.environ["PN2_PER_SCENE_SEED"] = "1"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
    with tempfile.TemporaryDirectory() as d:
        root = sk.make_dataset(d, name="kitti", n_scenes=64, split="val", seed=1, npoints=22000, n_invisible=98000, alias_to=n)
        ds = mod.KittiRCNNDataset(root, npoints=16384, split="val", mode="EVAL", random_select=True, classes="Car")
        t = time.perf_counter()
        for i in range(64):
            ds[i]
        print("one process, native per-scene pipeline: %.2f ms per scene" % ((time.perf_counter() - t) / 64 * 1e3))
        for native, shared, workers in ((True, True, 4), (True, True, 8), (True, True, 16), (True, False, 4), (False, True, 4), (True, True, 4)):
            mod.NATIVE_DATAPATH, mod.SHARED_BATCHES = native, shared
            dl = DataLoader(ds, batch_size=16, shuffle=False, pin_memory=True, num_workers=workers, collate_fn=ds.collate_batch)
            t = time.perf_counter()
            total = 0.0
            for b in dl:
                total += float(torch.from_numpy(b["pts_input"]).cuda(non_blocking=True).float().sum())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            print("workers %2d  %s  %s: %7.1f scenes/s (%d scenes in %.2f s)" % (workers, "native" if native else "numpy ",
                                                                                 "shared-memory batches" if shared else "pickled batches     ", n / dt, n, dt))


if __name__ == "__main__":
    main()
