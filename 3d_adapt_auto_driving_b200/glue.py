"""Launch plumbing for csrc/glue.cu: the small stages between the big kernels of one inference step, one kernel each
instead of the 20-60 elementwise / gather / scatter launches the reference's torch statements become.

    decode_bbox          lib/utils/bbox_transform.py:24-121 (+ proposal_layer.py:23)
    proposal_select      lib/rpn/proposal_layer.py:58-100 (both depth bands, up to the NMS input)
    proposal_assemble    lib/rpn/proposal_layer.py:107-119
    rcnn_post_prepare    tools/eval_rcnn.py:516-535, 611-620
    rcnn_post_assemble   tools/eval_rcnn.py:621-627

The kernels execute the same IEEE single-precision operations in the same order as the torch statements (explicit
round-to-nearest intrinsics, no FMA contraction); tests/test_glue_gpu.py compares every output bit for bit with the
torch composition (bbox_transform.decode_bbox_target, ProposalLayer with fused=False, Detector.postprocess_torch).
ROT_MODE / ROT_MODE_POOL select how the K = 2 batched matmul of rotate_pc_along_y_torch rounds -- a property of the library
kernel torch dispatches to, which differs between the two call sites: for the (rows, 1, 2) x (rows, 2, 2) product of the
box decoding it is mode 2 = fl(fl(x * r0) + fl(z * r1)) (measured on B200: 100 % of the elements equal in mode 2, 98.9 % /
98.8 % in the FMA orders 0 / 1), for the (rois, 512, 2) x (rois, 2, 2) product of the canonical transform mode 0.
tests/test_glue_gpu.py pins both against torch on the box."""
import ctypes
import os

import torch

from . import cabi
from .cabi import i32, ptr

ROT_MODE = int(os.environ.get("PN2_ROT_MODE", "2"))
# the same product over the (rois, 512, 2) pooled points goes to a different library kernel, which does contract:
# fma(z, r1, fl(x * r0)) -- mode 0 (measured: 100 % equal in mode 0, 78 % / 84 % in modes 1 / 2)
ROT_MODE_POOL = int(os.environ.get("PN2_ROT_MODE_POOL", "0"))
ENABLED = os.environ.get("PN2_GLUE", "1") != "0"


def _f64(x):
    return ctypes.c_double(float(x))


_anchor_cache = {}


def _anchor_host(anchor_size):
    """the three mean sizes as a host float[3].  Callers on the hot path pass the numpy row of cfg.CLS_MEAN_SIZE (no
    device round trip); a CUDA tensor is read back once and cached (a .tolist() synchronises: not during graph capture)."""
    if not isinstance(anchor_size, torch.Tensor):
        return (ctypes.c_float * 3)(*[float(v) for v in anchor_size])
    key = (anchor_size.data_ptr(), str(anchor_size.device), anchor_size._version)
    got = _anchor_cache.get(key)
    if got is None:
        if len(_anchor_cache) > 16:
            _anchor_cache.clear()
        got = (ctypes.c_float * 3)(*anchor_size.detach().float().cpu().tolist())
        _anchor_cache[key] = got
    return got


def decode_bbox(roi, reg, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_xz_fine=True, get_y_by_bin=False,
                loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=False, y_bottom=False):
    """decode_bbox_target on the device in one launch.  roi (rows, 3 | 7), reg (rows, C) contiguous f32."""
    roi, reg = roi.contiguous(), reg.contiguous()
    rows, c = reg.shape
    out = torch.empty((rows, 7), dtype=torch.float32, device=reg.device)
    cabi.call("pn2_decode_bbox_f32", ptr(roi), i32(roi.shape[1]), ptr(reg), i32(c), ptr(out), ctypes.c_longlong(rows),
              _f64(loc_scope), _f64(loc_bin_size), i32(num_head_bin), _anchor_host(anchor_size), i32(bool(get_xz_fine)),
              i32(bool(get_y_by_bin)), _f64(loc_y_scope), _f64(loc_y_bin_size), i32(bool(get_ry_fine)), i32(bool(y_bottom)),
              i32(ROT_MODE))
    return out


def argsort_desc(scores):
    """torch.sort(scores, dim=1, descending=True)[1] in one launch (csrc/glue.cu); rows of up to 16384 scores."""
    B, N = scores.shape
    order = torch.empty((B, N), dtype=torch.int64, device=scores.device)
    cabi.call("pn2_argsort_desc_f32", ptr(scores.contiguous()), ptr(order), i32(B), i32(N))
    return order


def proposal_select(order, props, pre0, pre1):
    """order (B, N) int64 descending-score order, props (B, N, 7) -> cidx0, cidx1, bev0, bev1, cnt (2, B)."""
    B, N = order.shape
    dev = props.device
    cidx0 = torch.empty((B, pre0), dtype=torch.int32, device=dev)
    cidx1 = torch.empty((B, pre1), dtype=torch.int32, device=dev)
    bev0 = torch.empty((B, pre0, 5), dtype=torch.float32, device=dev)
    bev1 = torch.empty((B, pre1, 5), dtype=torch.float32, device=dev)
    cnt = torch.empty((2, B), dtype=torch.int32, device=dev)
    cabi.call("pn2_proposal_select_f32", ptr(order), ptr(props), i32(B), i32(N), i32(pre0), i32(pre1), ptr(cidx0),
              ptr(cidx1), ptr(bev0), ptr(bev1), ptr(cnt))
    return cidx0, cidx1, bev0, bev1, cnt


def nms_raw(bev, counts, thresh, rotated, max_keep):
    """pn2_nms_bev_f32 with un-initialised outputs (the assemble kernels read only the first num entries)."""
    P, n, _ = bev.shape
    keep = torch.empty((P, max(max_keep, 1)), dtype=torch.int64, device=bev.device)
    num = torch.empty((P,), dtype=torch.int32, device=bev.device)
    cabi.call("pn2_nms_bev_f32", ptr(bev), i32(P), i32(n), i32(n), ptr(counts), cabi.f32(thresh), i32(1 if rotated else 0),
              i32(max_keep), ptr(keep), ptr(num))
    return keep, num


def nms_raw_pair(bev0, counts0, max_keep0, bev1, counts1, max_keep1, thresh, rotated):
    """nms_raw for two sets of problems in one launch (pn2_nms_bev_pair_f32) -> (keep0, num0, keep1, num1)."""
    P0, n0, _ = bev0.shape
    P1, n1, _ = bev1.shape
    keep0 = torch.empty((P0, max(max_keep0, 1)), dtype=torch.int64, device=bev0.device)
    keep1 = torch.empty((P1, max(max_keep1, 1)), dtype=torch.int64, device=bev0.device)
    num0 = torch.empty((P0,), dtype=torch.int32, device=bev0.device)
    num1 = torch.empty((P1,), dtype=torch.int32, device=bev0.device)
    cabi.call("pn2_nms_bev_pair_f32", ptr(bev0), i32(P0), i32(n0), i32(n0), ptr(counts0), i32(max_keep0), ptr(keep0), ptr(num0),
              ptr(bev1), i32(P1), i32(n1), i32(n1), ptr(counts1), i32(max_keep1), ptr(keep1), ptr(num1), cabi.f32(thresh),
              i32(1 if rotated else 0))
    return keep0, num0, keep1, num1


def proposal_assemble(props, scores, cidx0, cidx1, keep0, keep1, num0, num1, post0, post1):
    B, N = scores.shape
    rois = torch.empty((B, post0 + post1, 7), dtype=torch.float32, device=props.device)
    roi_scores = torch.empty((B, post0 + post1), dtype=torch.float32, device=props.device)
    cabi.call("pn2_proposal_assemble_f32", ptr(props), ptr(scores), i32(B), i32(N), ptr(cidx0), ptr(cidx1),
              i32(cidx0.shape[1]), i32(cidx1.shape[1]), ptr(keep0), ptr(keep1), ptr(num0), ptr(num1), i32(post0), i32(post1),
              ptr(rois), ptr(roi_scores))
    return rois, roi_scores


def rcnn_post_prepare(rois, reg, cls, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_y_by_bin, loc_y_scope,
                      loc_y_bin_size, score_thresh):
    """rois (B, M, 7), reg (B * M, C), cls (B * M,) raw scores -> boxes_sorted (B, M, 7), scores_sorted (B, M),
    bev (B, M, 5), counts (B,) int32."""
    B, M, _ = rois.shape
    dev = rois.device
    rois, reg, cls = rois.contiguous(), reg.contiguous(), cls.contiguous()
    boxes = torch.empty((B, M, 7), dtype=torch.float32, device=dev)
    scores = torch.empty((B, M), dtype=torch.float32, device=dev)
    bev = torch.empty((B, M, 5), dtype=torch.float32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    cabi.call("pn2_rcnn_post_prepare_f32", ptr(rois), ptr(reg), i32(reg.shape[1]), ptr(cls), i32(B), i32(M),
              _f64(loc_scope), _f64(loc_bin_size), i32(num_head_bin), _anchor_host(anchor_size), i32(bool(get_y_by_bin)),
              _f64(loc_y_scope), _f64(loc_y_bin_size), _f64(score_thresh), i32(ROT_MODE), ptr(boxes), ptr(scores), ptr(bev),
              ptr(counts))
    return boxes, scores, bev, counts


def rcnn_post_assemble(boxes_sorted, scores_sorted, keep, num):
    B, M, _ = boxes_sorted.shape
    rec = torch.empty((B, M, 8), dtype=torch.float32, device=boxes_sorted.device)
    cabi.call("pn2_rcnn_post_assemble_f32", ptr(boxes_sorted), ptr(scores_sorted), ptr(keep), ptr(num), i32(B), i32(M),
              ptr(rec))
    return rec


def roipool_canonical(xyz, rois, extra_width, score_raw, score_thresh, depth, depth_norm, feats, sampled, off2):
    """rcnn_net.py:126-154 in one launch (csrc/roipool3d.cu: pn2_roipool3d_canon_f32) -> pooled (B * M, sampled, off2 + C),
    empty (B, M) int32."""
    B, N, C2 = feats.shape
    M = rois.shape[1]
    ld = off2 + C2
    pooled = torch.empty((B * M, sampled, ld), dtype=torch.float32, device=xyz.device)
    empty = torch.zeros((B, M), dtype=torch.int32, device=xyz.device)
    cabi.call("pn2_roipool3d_canon_f32", ptr(xyz.contiguous()), ptr(rois.contiguous()), _f64(extra_width),
              ptr(score_raw.contiguous()), _f64(score_thresh), ptr(depth.contiguous()), _f64(depth_norm),
              ptr(feats.contiguous()), i32(C2), ptr(pooled), i32(ld), i32(off2), ptr(empty), i32(B), i32(N), i32(M),
              i32(sampled), i32(ROT_MODE_POOL), work=4.0 * B * M * sampled * (ld + C2 + 5))
    return pooled, empty
