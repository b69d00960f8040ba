"""The STOCK reference network on a GPU (test / baseline infrastructure; needs a CUDA device).

pointrcnn/lib/net/{point_rcnn,rpn,rcnn_net,pointnet2_msg}.py, pointnet2_lib/pointnet2/{pointnet2_modules,
pointnet2_utils,pytorch_utils}.py, lib/rpn/proposal_layer.py and lib/utils/{bbox_transform,kitti_utils,
iou3d/iou3d_utils,roipool3d/roipool3d_utils}.py are imported UNMODIFIED -- from /root/reference in the build
container, from the git-ignored verbatim copy baseline/_ref/pointrcnn on the GPU box (stage_reference_tree(),
run by __graft_entry__.build()).  Only the three compiled extension modules are supplied here, because the
reference's .cpp wrappers cannot build against torch 2.11 (THC is gone):

  backend "legacy": `pointnet2_cuda`, `iou3d_cuda`, `roipool3d_cuda` whose functions pass the caller's device
      pointers to oracle/_ref/libpn2_legacy.so = the reference's .cu files compiled unchanged.  nms_gpu follows
      iou3d.cpp:73-121 (mask kernel, blocking D2H of the u64 matrix, greedy pass on the host, keep written into
      the caller's CPU LongTensor).  This is the bench's `--impl reference` arm and the golden for the GPU tests.
  backend "b200":   the same three module names bound to the package's ctypes stubs:
      the reference's own Python on the sm_100a kernels.

Nothing under 3d_adapt_auto_driving_b200/ imports this file.
"""
import contextlib
import ctypes
import os
import shutil
import sys
import types

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(os.environ.get("PN2_REFERENCE_ROOT", "/root/reference"), "pointrcnn")
STAGED = os.path.join(ROOT, "baseline", "_ref", "pointrcnn")
PKG = "3d_adapt_auto_driving_b200"
# bytes moved between host and device by the reference's eval loop (bench.py declares them): the cloud batch up; per NMS
# call the u64 suppression matrix down and the keep list up again (iou3d.cpp:92-98, iou3d_utils.py:69); detections down
COPIED = {"h2d": 0, "d2h": 0}


def stage_reference_tree():
    """verbatim, git-ignored copy of the reference's Python (no .cu/.cpp, no data) so that it travels to the GPU
    box with the snapshot: lib/, pointnet2_lib/pointnet2/*.py, tools/{eval_rcnn.py,_init_path.py,cfgs,train_utils}."""
    if not os.path.isdir(SRC):
        return STAGED if os.path.isdir(STAGED) else None
    keep = (".py", ".yaml")

    def copy_tree(rel):
        for d, _, files in os.walk(os.path.join(SRC, rel)):
            for f in files:
                if f.endswith(keep):
                    dst = os.path.join(STAGED, os.path.relpath(os.path.join(d, f), SRC))
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(d, f)):
                        shutil.copyfile(os.path.join(d, f), dst)

    for rel in ("lib", "pointnet2_lib/pointnet2", "tools/cfgs", "tools/train_utils"):
        copy_tree(rel)
    for f in ("eval_rcnn.py", "_init_path.py"):
        dst = os.path.join(STAGED, "tools", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, "tools", f), dst)
    return STAGED


def ref_root():
    if os.path.isdir(SRC):
        return SRC
    if os.path.isdir(STAGED):
        return STAGED
    return None


def available(backend="legacy"):
    if ref_root() is None or not torch.cuda.is_available():
        return False
    if backend == "legacy":
        from . import legacy
        return legacy.available()
    return True


class _AttrDict(dict):
    """stand-in for easydict.EasyDict (absent offline)."""

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


def _p(t):
    assert t.is_cuda and t.is_contiguous(), "extension stubs take contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _legacy_stubs():
    from . import legacy, oracle as orc
    L = legacy.lib()
    greedy = orc.lib().orc_nms_greedy_from_mask
    greedy.restype = ctypes.c_int
    f32 = ctypes.c_float

    def s():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    p2 = types.ModuleType("pointnet2_cuda")

    def fps_w(b, n, m, points, temp, idx):
        L.legacy_fps(b, n, m, _p(points), _p(temp), _p(idx), s()); return 1

    def gather_w(b, c, n, npoints, points, idx, out):
        L.legacy_gather(b, c, n, npoints, _p(points), _p(idx), _p(out), s()); return 1

    def bq_w(b, n, m, radius, nsample, new_xyz, xyz, idx):
        L.legacy_ball_query(b, n, m, f32(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), s()); return 1

    def group_w(b, c, n, npoints, nsample, points, idx, out):
        L.legacy_group(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), s()); return 1

    def nn_w(b, n, m, unknown, known, dist2, idx):
        L.legacy_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), s())

    def interp_w(b, c, m, n, points, idx, weight, out):
        L.legacy_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), s())

    p2.furthest_point_sampling_wrapper, p2.gather_points_wrapper, p2.ball_query_wrapper = fps_w, gather_w, bq_w
    p2.group_points_wrapper, p2.three_nn_wrapper, p2.three_interpolate_wrapper = group_w, nn_w, interp_w

    iou = types.ModuleType("iou3d_cuda")

    def nms_w(normal):
        def f(boxes, keep, thresh):
            """iou3d.cpp:73-121 / :124-169: the launcher works on the legacy default stream."""
            n = boxes.size(0)
            if n == 0:
                return 0
            mask = torch.zeros((n, (n + 63) // 64), dtype=torch.int64, device=boxes.device)
            torch.cuda.current_stream().synchronize()
            (L.legacy_nms_normal_mask if normal else L.legacy_nms_mask)(_p(boxes), _p(mask), n, f32(thresh))
            host = mask.cpu().numpy()                        # the reference's blocking cudaMemcpy D2H
            k = greedy(host.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_void_p(keep.data_ptr()))
            COPIED["d2h"] += host.nbytes
            COPIED["h2d"] += 8 * int(k)
            return int(k)
        return f

    def pair_w(fn):
        def f(a, b, out):
            torch.cuda.current_stream().synchronize()
            fn(a.size(0), _p(a), b.size(0), _p(b), _p(out))
            torch.cuda.synchronize()
            return 1
        return f

    iou.nms_gpu, iou.nms_normal_gpu = nms_w(False), nms_w(True)
    iou.boxes_overlap_bev_gpu, iou.boxes_iou_bev_gpu = pair_w(L.legacy_boxes_overlap_bev), pair_w(L.legacy_boxes_iou_bev)

    rp = types.ModuleType("roipool3d_cuda")

    def roipool_fw(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
        # roipool3d.cpp:17-45; the reference launcher cudaMallocs / cudaFrees its flag buffer on the legacy stream
        torch.cuda.current_stream().synchronize()
        L.legacy_roipool3d(xyz.size(0), xyz.size(1), boxes3d.size(1), pts_feature.size(2), pooled_features.size(2),
                           _p(xyz), _p(boxes3d), _p(pts_feature), _p(pooled_features), _p(pooled_empty_flag))
        return 1

    rp.forward = roipool_fw
    return {"pointnet2_cuda": p2, "iou3d_cuda": iou, "roipool3d_cuda": rp}


def _b200_stubs():
    import importlib
    return {name: importlib.import_module(PKG + "." + name) for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda")}


_MODULE_PREFIXES = ("lib", "pointnet2_lib", "pointnet2_msg", "train_utils", "tools")


@contextlib.contextmanager
def reference_imports(backend="legacy"):
    """Inside the block the reference's packages are importable with the chosen extension stubs; afterwards its
    modules are dropped from sys.modules (objects created inside keep working, they hold their globals)."""
    root = ref_root()
    if root is None:
        raise RuntimeError("no reference tree: neither %s nor %s exists" % (SRC, STAGED))
    before = dict(sys.modules)
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    stubs = dict(_legacy_stubs() if backend == "legacy" else _b200_stubs(), easydict=ed)
    sys.modules.update(stubs)
    paths = [root, os.path.join(root, "lib", "net"), os.path.join(root, "tools")]
    for p in paths:
        sys.path.insert(0, p)
    old_load = yaml.load
    yaml.load = lambda f, *a, **k: old_load(f, Loader=yaml.SafeLoader)       # config.py:cfg_from_file, PyYAML >= 6
    try:
        yield root
    finally:
        yaml.load = old_load
        for p in paths:
            sys.path.remove(p)
        for k in list(sys.modules):
            if k not in before and (k in stubs or k.split(".")[0] in _MODULE_PREFIXES):
                del sys.modules[k]
        for k in stubs:
            if k in before:
                sys.modules[k] = before[k]
            else:
                sys.modules.pop(k, None)


class Reference:
    """The reference PointRCNN (TEST mode, default.yaml, `--eval_mode rcnn`: eval_rcnn.py:860-866) on `device`, and
    the per-batch body of eval_one_epoch_joint (eval_rcnn.py:497-535, 611-627) built from the reference's own
    decode_bbox_target / boxes3d_to_bev_torch / iou3d_utils.nms_gpu."""

    def __init__(self, state_dict, device, backend="legacy"):
        self.backend = backend
        self.device = torch.device(device)
        with reference_imports(backend) as root, torch.cuda.device(self.device):
            from lib.config import cfg, cfg_from_file
            cfg_from_file(os.path.join(root, "tools", "cfgs", "default.yaml"))
            cfg.TAG = "default"
            cfg.RCNN.ENABLED = True
            cfg.RPN.ENABLED = cfg.RPN.FIXED = True
            from lib.net.point_rcnn import PointRCNN
            from lib.utils.bbox_transform import decode_bbox_target
            import lib.utils.kitti_utils as kitti_utils
            import lib.utils.iou3d.iou3d_utils as iou3d_utils
            self.cfg, self.decode, self.kitti_utils, self.iou3d_utils = cfg, decode_bbox_target, kitti_utils, iou3d_utils
            model = PointRCNN(num_classes=2, use_xyz=True, mode="TEST")
        model.load_state_dict(state_dict, strict=True)
        self.model = model.cuda(self.device).eval()
        self.mean_size = torch.from_numpy(cfg.CLS_MEAN_SIZE[0]).cuda(self.device)

    @torch.no_grad()
    def forward(self, pts_input):
        """model({'pts_input': (B,N,3) cuda}) -> dict of tensors (eval_rcnn.py:497-500)."""
        with torch.cuda.device(self.device):
            out = self.model({"pts_input": pts_input})
        return {k: v for k, v in out.items() if isinstance(v, torch.Tensor)}

    @torch.no_grad()
    def eval_batch(self, pts_host):
        """One iteration of the eval loop for a host batch (B,N,3): H2D, forward, decode, score threshold, per-scene
        rotated NMS, detections copied to the host -> list over scenes of (boxes3d (k,7), raw scores (k,)) ndarrays
        (scenes without a box above the threshold give empty arrays, where the reference `continue`s)."""
        cfg = self.cfg
        with torch.cuda.device(self.device):
            inputs = pts_host.cuda(non_blocking=True).float()                                   # :498
            COPIED["h2d"] += pts_host.numel() * pts_host.element_size()
            ret_dict = self.model({"pts_input": inputs})
            batch_size = inputs.shape[0]
            roi_boxes3d = ret_dict["rois"]
            rcnn_cls = ret_dict["rcnn_cls"].view(batch_size, -1, ret_dict["rcnn_cls"].shape[1])
            rcnn_reg = ret_dict["rcnn_reg"].view(batch_size, -1, ret_dict["rcnn_reg"].shape[1])
            pred_boxes3d = self.decode(roi_boxes3d.view(-1, 7), rcnn_reg.view(-1, rcnn_reg.shape[-1]),
                                       anchor_size=self.mean_size, loc_scope=cfg.RCNN.LOC_SCOPE,
                                       loc_bin_size=cfg.RCNN.LOC_BIN_SIZE, num_head_bin=cfg.RCNN.NUM_HEAD_BIN,
                                       get_xz_fine=True, get_y_by_bin=cfg.RCNN.LOC_Y_BY_BIN,
                                       loc_y_scope=cfg.RCNN.LOC_Y_SCOPE, loc_y_bin_size=cfg.RCNN.LOC_Y_BIN_SIZE,
                                       get_ry_fine=True).view(batch_size, -1, 7)                  # :516-524
            raw_scores = rcnn_cls                                                               # :527-531
            norm_scores = torch.sigmoid(raw_scores)
            inds = norm_scores > cfg.RCNN.SCORE_THRESH                                          # :612
            out = []
            for k in range(batch_size):
                cur_inds = inds[k].view(-1)
                if cur_inds.sum() == 0:
                    out.append((torch.zeros((0, 7)).numpy(), torch.zeros((0,)).numpy()))
                    continue
                boxes_sel = pred_boxes3d[k, cur_inds]
                raw_sel = raw_scores[k, cur_inds]
                bev = self.kitti_utils.boxes3d_to_bev_torch(boxes_sel)
                keep = self.iou3d_utils.nms_gpu(bev, raw_sel, cfg.RCNN.NMS_THRESH).view(-1)         # :620-621
                out.append((boxes_sel[keep].cpu().numpy(), raw_sel[keep].view(-1).cpu().numpy()))  # :622-624
                COPIED["d2h"] += out[-1][0].nbytes + out[-1][1].nbytes
        return out
