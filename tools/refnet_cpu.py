"""Run the REFERENCE's own network code on the CPU (build container only; test infrastructure).

pointrcnn/lib/net/{point_rcnn,rpn,rcnn_net,pointnet2_msg}.py, pointnet2_lib/pointnet2/{pointnet2_modules,
pointnet2_utils,pytorch_utils}.py, lib/rpn/proposal_layer.py, lib/utils/{bbox_transform,kitti_utils,iou3d/iou3d_utils,
roipool3d/roipool3d_utils}.py are imported UNMODIFIED from /root/reference.  What cannot exist here is replaced at the
module boundary, nothing inside the reference's Python is touched:
  * the three CUDA extensions (`pointnet2_cuda`, `iou3d_cuda`, `roipool3d_cuda`) become stub modules with the same
    function signatures (pointnet2_api.cpp:10-24, iou3d.cpp:180-186, roipool3d.cpp:107-109) whose bodies are the C
    restatements of the reference kernels in oracle/ (each pinned to the reference's kernels' outputs,
    tests/test_golden_cpu.py), writing into the caller-allocated CPU tensors;
  * torch.cuda.{Float,Int,Long}Tensor construct CPU tensors and Tensor.cuda() is the identity while the network runs;
  * `easydict` (absent) is a small attribute dict, yaml.load gets the Loader argument PyYAML 6 demands.
Purpose: oracle/cpu_forward.py -- the CPU port every GPU parity test and the bench's cpu_baseline lean on -- restates
the reference's COMPOSITION (module order, tensor layouts, proposal layer, post-processing).  With this harness the
restatement is checked against the reference's own composition code on the same weights and inputs
(tests/test_refnet_vs_port_cpu.py), and golden vectors are written for the GPU box (tools/make_refnet_fixture.py).
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pointrcnn"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available():
    return os.path.isdir(REF)


class _AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _np(t):
    return t.detach().contiguous().numpy()


def _put(dst, arr):
    dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view_as(dst))


def _extension_stubs():
    from oracle import oracle as orc
    orc.lib()
    p2 = types.ModuleType("pointnet2_cuda")

    def fps_w(b, n, m, xyz, temp, out):
        idx, t = orc.fps(_np(xyz), m, temp=_np(temp))
        _put(out, idx); _put(temp, t)
        return 1

    def gather_w(b, c, n, npoints, points, idx, out):
        _put(out, orc.gather_points(_np(points), _np(idx)))
        return 1

    def bq_w(b, n, m, radius, nsample, new_xyz, xyz, idx):
        _put(idx, orc.ball_query(radius, nsample, _np(xyz), _np(new_xyz)))
        return 1

    def group_w(b, c, n, npoints, nsample, points, idx, out):
        _put(out, orc.group_points(_np(points), _np(idx)))
        return 1

    def nn_w(b, n, m, unknown, known, dist2, idx):
        d2, i = orc.three_nn(_np(unknown), _np(known))
        _put(dist2, d2); _put(idx, i)

    def interp_w(b, c, m, n, points, idx, weight, out):
        _put(out, orc.three_interpolate(_np(points), _np(idx), _np(weight)))

    p2.furthest_point_sampling_wrapper, p2.gather_points_wrapper, p2.ball_query_wrapper = fps_w, gather_w, bq_w
    p2.group_points_wrapper, p2.three_nn_wrapper, p2.three_interpolate_wrapper = group_w, nn_w, interp_w

    iou = types.ModuleType("iou3d_cuda")

    def nms_w(rotated):
        def f(boxes, keep, thresh):
            k = orc.nms_rotated(_np(boxes), thresh) if rotated else orc.nms_normal(_np(boxes), thresh)
            keep[:len(k)] = torch.from_numpy(k)
            return len(k)
        return f

    iou.nms_gpu, iou.nms_normal_gpu = nms_w(True), nms_w(False)
    iou.boxes_overlap_bev_gpu = lambda a, b, out: _put(out, orc.boxes_overlap_bev(_np(a), _np(b)))
    iou.boxes_iou_bev_gpu = lambda a, b, out: _put(out, orc.boxes_overlap_bev(_np(a), _np(b), iou=True))

    rp = types.ModuleType("roipool3d_cuda")

    def roipool_fw(pts, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
        pooled, empty = orc.roipool3d(_np(pts), _np(pts_feature), _np(boxes3d), sampled=pooled_features.shape[2])
        _put(pooled_features, pooled); _put(pooled_empty_flag, empty)
        return 1

    rp.forward = roipool_fw
    return {"pointnet2_cuda": p2, "iou3d_cuda": iou, "roipool3d_cuda": rp}


@contextlib.contextmanager
def cpu_cuda():
    """torch.cuda tensor constructors build CPU tensors, .cuda() is the identity and get_device() names the CPU
    (bbox_transform.py:40 does anchor_size.to(roi_box3d.get_device())) inside the block."""
    saved = (torch.cuda.FloatTensor, torch.cuda.IntTensor, torch.cuda.LongTensor, torch.Tensor.cuda, torch.Tensor.get_device)
    torch.cuda.FloatTensor, torch.cuda.IntTensor, torch.cuda.LongTensor = torch.FloatTensor, torch.IntTensor, torch.LongTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.get_device = lambda self: torch.device("cpu")
    try:
        yield
    finally:
        (torch.cuda.FloatTensor, torch.cuda.IntTensor, torch.cuda.LongTensor, torch.Tensor.cuda,
         torch.Tensor.get_device) = saved


@contextlib.contextmanager
def reference_imports():
    """Inside the block the reference's packages are importable (stubs in place, its paths first on sys.path);
    afterwards its `lib` / `pointnet2_lib` modules are removed from sys.modules again so that they cannot shadow
    anything else in the process.  Objects created inside keep working."""
    before = set(sys.modules)
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    stubs = dict(_extension_stubs(), easydict=ed)
    shadowed = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    paths = [REF, os.path.join(REF, "lib", "net")]
    for p in paths:
        sys.path.insert(0, p)
    old_load = yaml.load
    yaml.load = lambda f, *a, **k: old_load(f, Loader=yaml.SafeLoader)
    try:
        with cpu_cuda():
            yield
    finally:
        yaml.load = old_load
        for p in paths:
            sys.path.remove(p)
        for k in set(sys.modules) - before:
            if k == "lib" or k.startswith("lib.") or k.startswith("pointnet2_lib") or k == "pointnet2_msg" or k in stubs:
                del sys.modules[k]
        for k, v in shadowed.items():
            if v is not None:
                sys.modules[k] = v


def build_reference_model(state_dict):
    """The reference PointRCNN (TEST mode), configured like `eval_rcnn.py --cfg_file cfgs/default.yaml --eval_mode rcnn`
    (eval_rcnn.py:860-866), with the given parameters, in eval()."""
    with reference_imports():
        from lib.config import cfg, cfg_from_file
        cfg_from_file(os.path.join(REF, "tools", "cfgs", "default.yaml"))
        cfg.TAG = "default"
        cfg.RCNN.ENABLED = True
        cfg.RPN.ENABLED = cfg.RPN.FIXED = True
        from lib.net.point_rcnn import PointRCNN
        model = PointRCNN(num_classes=2, use_xyz=True, mode="TEST")
    model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model


def reference_forward(model, pts_input):
    """model({'pts_input': (B,N,3)}) as eval_one_epoch_joint calls it (eval_rcnn.py:497-500) -> dict of CPU tensors."""
    with cpu_cuda(), torch.no_grad():
        out = model({"pts_input": pts_input})
    return {k: v for k, v in out.items() if isinstance(v, torch.Tensor)}


# ---- the reference's eval_rcnn.py itself, end to end on the CPU ---------------------------------------------------------
_TENSORBOARDX = "class SummaryWriter(object):\n    def __init__(self, *a, **k): pass\n    def add_scalar(self, *a, **k): pass\n"


def stage_reference_tree(dest):
    """<dest>/pointrcnn with `lib`, `pointnet2_lib`, `tools/train_utils`, `tools/cfgs` SYMLINKED to /root/reference and
    tools/{eval_rcnn.py,_init_path.py} copied there at run time (the script derives its data root from its own
    realpath, eval_rcnn.py:854, and the reference tree is read-only) -> the tools directory to run in."""
    import shutil
    root = os.path.join(dest, "pointrcnn")
    tools = os.path.join(root, "tools")
    os.makedirs(tools)
    for name in ("lib", "pointnet2_lib"):
        os.symlink(os.path.join(REF, name), os.path.join(root, name))
    for name in ("train_utils", "cfgs"):
        os.symlink(os.path.join(REF, "tools", name), os.path.join(tools, name))
    for name in ("eval_rcnn.py", "_init_path.py"):
        shutil.copyfile(os.path.join(REF, "tools", name), os.path.join(tools, name))
    os.makedirs(os.path.join(tools, "tensorboardX"))
    with open(os.path.join(tools, "tensorboardX", "__init__.py"), "w") as f:
        f.write(_TENSORBOARDX)
    return tools


def run_reference_eval(tools_dir, argv, timeout=1800):
    """`python eval_rcnn.py <argv>` in tools_dir, on the CPU: this file is the interpreter's entry point, installs the
    stubs and then executes the script as __main__."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--run", "eval_rcnn.py"] + list(argv)
    return subprocess.run(cmd, cwd=tools_dir, capture_output=True, text=True, timeout=timeout)


def _main_run(script, argv):
    import runpy
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    sys.modules.update(dict(_extension_stubs(), easydict=ed))
    old_load = yaml.load
    yaml.load = lambda f, *a, **k: old_load(f, Loader=yaml.SafeLoader)
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.argv = [script] + argv
    sys.path.insert(0, os.getcwd())
    import _init_path  # noqa: F401  (the script's own path set-up, executed a little earlier)
    # The one accommodation: eval_rcnn.py:862 passes far_points= to a constructor whose parameter is called
    # npoints_faraway (kitti_rcnn_dataset.py:13-16) -- a TypeError of the reference against itself.
    import lib.datasets.kitti_rcnn_dataset as ds_mod
    ref_init = ds_mod.KittiRCNNDataset.__init__

    def init_accepting_far_points(self, *a, far_points=None, **k):
        if far_points is not None:
            k["npoints_faraway"] = far_points
        ref_init(self, *a, **k)

    ds_mod.KittiRCNNDataset.__init__ = init_accepting_far_points
    with cpu_cuda():
        runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        _main_run(sys.argv[2], sys.argv[3:])
    else:
        raise SystemExit("usage: refnet_cpu.py --run <script.py> [script args]")
