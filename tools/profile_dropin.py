#!/usr/bin/env python
"""Where does the main process of the UNMODIFIED eval_rcnn.py spend its time when it runs on this package?  Stages the
drop-in tree + a synthetic data set (tools/run_config5.py's), runs the script in-process under cProfile and prints the top
functions by cumulative and by own time.   python tools/profile_dropin.py [--scenes 640] [--workers 4]"""
import argparse
import cProfile
import importlib
import io
import os
import pstats
import runpy
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=640)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--work", default="/tmp/pn2_profile_dropin")
    args = ap.parse_args()
    import torch
    sk = importlib.import_module(PKG + ".synthetic_kitti")
    et = importlib.import_module(PKG + ".evaltree")
    inf = importlib.import_module(PKG + ".inference")
    tu = importlib.import_module(PKG + ".train_utils")
    shutil.rmtree(args.work, ignore_errors=True)
    tree = et.make_eval_tree(os.path.join(args.work, "tree"), os.path.join(ROOT, "oracle", "_ref", "eval_rcnn.py"))
    sk.make_dataset(tree, name="kitti", n_scenes=64, split="val", seed=666, npoints=22000, n_invisible=98000, alias_to=args.scenes)
    model = inf.build_model(seed=0, device="cpu")
    with torch.no_grad():
        model.rcnn_net.cls_layer[-1].conv.bias.fill_(1.0)
    ckpt = os.path.join(args.work, "checkpoint_epoch_1")
    tu.save_checkpoint(tu.checkpoint_state(model, None, 1, 1), filename=ckpt)
    tools = os.path.join(tree, "tools")
    os.chdir(tools)
    sys.path.insert(0, tools)
    os.environ["PN2_PER_SCENE_SEED"] = "1"
    sys.argv = ["eval_rcnn.py", "--cfg_file", "cfgs/default.yaml", "--eval_mode", "rcnn", "--ckpt", ckpt + ".pth", "--batch_size", "16",
                "--workers", str(args.workers), "--output_dir", os.path.join(args.work, "out")]
    prof = cProfile.Profile()
    prof.enable()
    try:
        runpy.run_path(os.path.join(tools, "eval_rcnn.py"), run_name="__main__")
    finally:
        prof.disable()
    for key, n in (("cumulative", 45), ("tottime", 30)):
        out = io.StringIO()
        pstats.Stats(prof, stream=out).strip_dirs().sort_stats(key).print_stats(n)
        print(out.getvalue()[:9000])


if __name__ == "__main__":
    main()
