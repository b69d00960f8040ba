"""Golden vectors for the host-side helpers between the network and the result files, produced by the REFERENCE
modules themselves (pointrcnn/lib/utils/{kitti_utils,calibration,bbox_transform,object3d}.py imported from
/root/reference, CPU only): box corners, image projections, bin-based box decoding in the RPN / RCNN configurations,
label parsing.  Run in the build container:  python tools/make_host_utils_fixture.py  ->  tests/golden/host_utils.npz.
tests/test_host_utils_cpu.py checks the mirrors against it bit for bit (and, where /root/reference exists, against
the live modules on fresh random inputs).

decode_bbox_target calls anchor_size.to(roi.get_device()), which raises on a CPU tensor; the anchor is handed over in
a wrapper whose .to() returns the CPU tensor -- the decoding arithmetic is the reference's, unmodified."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pointrcnn"
MEAN_SIZE = [1.52563191462, 1.62856739989, 3.88311640418]          # cfg.CLS_MEAN_SIZE[0] of the reference's default.yaml

# (loc_scope, loc_bin_size, num_head_bin, get_xz_fine, get_y_by_bin, loc_y_scope, loc_y_bin_size, get_ry_fine, roi is a box)
DECODE_CASES = [(3.0, 0.5, 12, True, False, 0.5, 0.25, False, False),      # RPN proposal layer
                (1.5, 0.5, 9, True, False, 0.5, 0.25, True, True),         # RCNN refinement
                (1.5, 0.5, 9, False, True, 0.5, 0.25, True, True),
                (3.0, 0.5, 12, True, True, 0.5, 0.25, False, False)]

LABEL_LINES = ["Car 0.00 0 -1.58 587.01 173.33 614.12 200.12 1.65 1.67 3.64 -0.65 1.71 46.70 -1.59",
               "Pedestrian 0.30 2 0.2 1.0 2.0 30.0 80.0 1.7 0.6 0.8 3.0 1.5 12.0 0.3 0.77",
               "DontCare -1 -1 -10 503.89 169.71 590.61 190.13 -1 -1 -1 -1000 -1000 -1000 -10",
               "Cyclist 0.60 3 1.0 1.0 2.0 30.0 20.0 1.7 0.6 1.8 3.0 1.5 12.0 0.3",
               "Van 0.10 1 1.0 1.0 2.0 30.0 30.0 1.7 0.6 1.8 3.0 1.5 12.0 0.3"]


class OnCpu:
    def __init__(self, t):
        self.t = t

    def to(self, device):
        return self.t


def reference_modules():
    sys.path.insert(0, REF)
    try:
        import lib.utils.kitti_utils as ku
        import lib.utils.calibration as cal
        import lib.utils.bbox_transform as bt
    finally:
        sys.path.remove(REF)
    return ku, cal, bt


def calib_file(dirname):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load
    path = os.path.join(dirname, "calib.txt")
    with open(path, "w") as f:
        f.write("\n".join(load("synthetic_kitti").CALIB_LINES) + "\n")
    return path


def inputs(seed, n_boxes=64, n_pts=512, n_rows=256):
    rs = np.random.RandomState(seed)
    boxes = np.concatenate([rs.uniform(-30, 30, (n_boxes, 1)), rs.uniform(-1, 3, (n_boxes, 1)), rs.uniform(-5, 70, (n_boxes, 1)),
                            rs.uniform(1, 4, (n_boxes, 3)), rs.uniform(-4, 4, (n_boxes, 1))], 1).astype(np.float32)
    pts = rs.uniform(-40, 70, (n_pts, 3)).astype(np.float32)
    dec = []
    for (scope, bsz, nh, fine, ybin, yscope, ybs, ryfine, roi7) in DECODE_CASES:
        nb, nby = int(scope / bsz) * 2, int(yscope / ybs) * 2
        c = nb * 2 + (nb * 2 if fine else 0) + (nby * 2 if ybin else 1) + nh * 2 + 3
        pred = rs.randn(n_rows, c).astype(np.float32)
        roi = boxes[rs.randint(0, n_boxes, n_rows)] if roi7 else (rs.randn(n_rows, 3) * 20).astype(np.float32)
        dec.append((pred, roi))
    return boxes, pts, dec


def evaluate(ku, cal, bt, calib_path, boxes, pts, dec, anchor):
    """The same calls on either set of modules -> dict of arrays."""
    out = {}
    calib = cal.Calibration(calib_path)
    out["rect"] = calib.lidar_to_rect(pts)
    out["img"], out["depth"] = calib.rect_to_img(out["rect"])
    for rot in (True, False):
        out["corners_%d" % rot] = ku.boxes3d_to_corners3d(boxes, rot)
    out["img_boxes"], out["img_corners"] = calib.corners3d_to_img_boxes(out["corners_1"])
    tb = torch.from_numpy(boxes)
    out["bev"] = ku.boxes3d_to_bev_torch(tb).numpy()
    out["enlarged"] = ku.enlarge_box3d(tb.clone(), 1.0).numpy()
    for i, ((scope, bsz, nh, fine, ybin, yscope, ybs, ryfine, _), (pred, roi)) in enumerate(zip(DECODE_CASES, dec)):
        out["decode_%d" % i] = bt.decode_bbox_target(torch.from_numpy(roi).clone(), torch.from_numpy(pred).clone(), scope, bsz,
                                                     nh, anchor, get_xz_fine=fine, get_y_by_bin=ybin, loc_y_scope=yscope,
                                                     loc_y_bin_size=ybs, get_ry_fine=ryfine).numpy()
    return out


def main():
    ku, cal, bt = reference_modules()
    with tempfile.TemporaryDirectory() as d:
        boxes, pts, dec = inputs(0)
        out = evaluate(ku, cal, bt, calib_file(d), boxes, pts, dec, OnCpu(torch.tensor(MEAN_SIZE)))
        lf = os.path.join(d, "label.txt")
        with open(lf, "w") as f:
            f.write("\n".join(LABEL_LINES) + "\n")
        objs = ku.get_objects_from_label(lf)
        out["obj_boxes3d"] = ku.objs_to_boxes3d(objs)
        out["obj_level"] = np.array([o.level for o in objs], np.int64)
        out["obj_text"] = np.array([o.to_kitti_format() for o in objs])
    path = os.path.join(ROOT, "tests", "golden", "host_utils.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
