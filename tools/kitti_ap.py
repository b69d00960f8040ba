"""KITTI AP of a result directory against a label directory, the core of the reference's evaluate/evaluate.py
(label readers -> get_official_eval_result) on this package's evaluator (evaluate/eval2.py: rotated overlaps on the
GPU, matching passes native).

    python tools/kitti_ap.py --label_dir <KITTI/object/training/label_2> --result_dir <.../final_result/data>
                             [--split_file <ImageSets/val.txt>] [--dataset kitti] [--classes Car]
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def evaluate_dirs(label_dir, result_dir, ids=None, dataset="kitti", classes=("Car",)):
    """-> (result text, ap dict, seconds spent reading, seconds spent evaluating); ids None = every result file."""
    kc = importlib.import_module(PKG + ".evaluate.kitti_common")
    ev = importlib.import_module(PKG + ".evaluate.eval2")
    t0 = time.time()
    dt_annos = kc.get_label_annos(result_dir, ids)
    if ids is None:
        ids = sorted(int(f[:-4]) for f in os.listdir(result_dir) if f.endswith(".txt") and len(f) == 10)
    gt_annos = kc.get_label_annos(label_dir, ids)
    t1 = time.time()
    result, ret = ev.get_official_eval_result(gt_annos, dt_annos, list(classes) if not isinstance(classes, int) else classes,
                                              dataset)
    return result, ret, t1 - t0, time.time() - t1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label_dir", required=True)
    ap.add_argument("--result_dir", required=True)
    ap.add_argument("--split_file", default=None, help="image ids to evaluate (default: every result file)")
    ap.add_argument("--dataset", default="kitti", choices=["kitti", "argo", "nusc", "lyft", "waymo"])
    ap.add_argument("--classes", default="Car")
    args = ap.parse_args()
    ids = [int(line) for line in open(args.split_file).read().split()] if args.split_file else None
    result, ret, t_read, t_eval = evaluate_dirs(args.label_dir, args.result_dir, ids, args.dataset, args.classes.split(","))
    print(result)
    print(json.dumps({k: float(v) for k, v in ret.items() if k != "result"}))
    print("read in %.2f s, evaluated in %.2f s" % (t_read, t_eval), file=sys.stderr)


if __name__ == "__main__":
    main()
