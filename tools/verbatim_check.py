#!/usr/bin/env python
"""Line-level similarity of a repo file with its reference counterpart, the way the round-1 review measured it:
stripped code lines longer than 12 characters (comments and docstring-only lines dropped) that also occur verbatim in the
reference file.   python tools/verbatim_check.py [repo_file reference_file] ...   (no arguments: the known host mirrors)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAIRS = [
    ("3d_adapt_auto_driving_b200/datasets/kitti_rcnn_dataset.py", "pointrcnn/lib/datasets/kitti_rcnn_dataset.py"),
    ("3d_adapt_auto_driving_b200/datasets/kitti_dataset.py", "pointrcnn/lib/datasets/kitti_dataset.py"),
    ("3d_adapt_auto_driving_b200/evaluate/eval2.py", "evaluate/eval2.py"),
    ("3d_adapt_auto_driving_b200/evaluate/kitti_common.py", "evaluate/kitti_common.py"),
    ("3d_adapt_auto_driving_b200/stat_norm/norm.py", "stat_norm/norm.py"),
    ("3d_adapt_auto_driving_b200/proposal_layer.py", "pointrcnn/lib/rpn/proposal_layer.py"),
    ("3d_adapt_auto_driving_b200/bbox_transform.py", "pointrcnn/lib/utils/bbox_transform.py"),
    ("3d_adapt_auto_driving_b200/pointnet2_modules.py", "pointrcnn/pointnet2_lib/pointnet2/pointnet2_modules.py"),
    ("3d_adapt_auto_driving_b200/net/rcnn_net.py", "pointrcnn/lib/net/rcnn_net.py"),
    ("3d_adapt_auto_driving_b200/net/rpn.py", "pointrcnn/lib/net/rpn.py"),
    ("3d_adapt_auto_driving_b200/kitti_utils.py", "pointrcnn/lib/utils/kitti_utils.py"),
    ("3d_adapt_auto_driving_b200/calibration.py", "pointrcnn/lib/utils/calibration.py"),
]


def code_lines(path):
    out = []
    for line in open(path, errors="replace"):
        s = line.strip()
        if len(s) <= 12 or s.startswith("#") or s.startswith('"""') or s.startswith("'''"):
            continue
        out.append(s)
    return out


def main():
    args = sys.argv[1:]
    pairs = list(zip(args[0::2], args[1::2])) if args else [(os.path.join(ROOT, a), os.path.join(REF, b)) for a, b in PAIRS]
    for mine, ref in pairs:
        if not (os.path.exists(mine) and os.path.exists(ref)):
            print("%-70s (missing)" % os.path.relpath(mine, ROOT))
            continue
        ref_set = set(code_lines(ref))
        lines = code_lines(mine)
        hit = sum(1 for s in lines if s in ref_set)
        print("%-70s %4d / %4d lines verbatim (%.0f %%)" % (os.path.relpath(mine, ROOT), hit, len(lines), 100.0 * hit / max(len(lines), 1)))


if __name__ == "__main__":
    main()
