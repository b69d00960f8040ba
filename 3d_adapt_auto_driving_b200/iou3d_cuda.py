"""Drop-in for the reference's `iou3d_cuda` extension (pointrcnn/lib/utils/iou3d/src/iou3d.cpp:174-179):
boxes_overlap_bev_gpu, boxes_iou_bev_gpu, nms_gpu, nms_normal_gpu with the same arguments.
nms_* keep the reference contract (boxes sorted by score on the device, `keep` a CPU
LongTensor that is filled, return value = number kept) by running the device-resident NMS
(pn2_nms_bev_f32) and copying the result into `keep`; the fused pipeline calls
iou3d_utils.nms_*_device instead and never leaves the GPU."""
import torch

from . import cabi
from .cabi import i32, f32, ptr


def _boxes(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise cabi.Pn2Error("%s must be a contiguous CUDA float32 tensor" % name)
    # the reference extension only looks at size(0) and the data pointer (iou3d.cpp:73-85), and the reference's own eval
    # loop hands it (N, 1, 5) boxes (tools/eval_rcnn.py:614-621: scores of shape (N, 1) index the boxes): accept any
    # contiguous layout of N x 5 floats
    if t.dim() < 2 or t.numel() != t.size(0) * 5:
        raise cabi.Pn2Error("%s must hold N x 5 floats, (N, 5) or (N, 1, 5)" % name)
    return t.view(t.size(0), 5)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    boxes_a, boxes_b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    cabi.call("pn2_boxes_overlap_bev_f32", ptr(boxes_a), i32(boxes_a.size(0)), ptr(boxes_b), i32(boxes_b.size(0)), ptr(ans_overlap))
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    boxes_a, boxes_b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    cabi.call("pn2_boxes_iou_bev_f32", ptr(boxes_a), i32(boxes_a.size(0)), ptr(boxes_b), i32(boxes_b.size(0)), ptr(ans_iou))
    return 1


def nms_device(boxes, thresh, rotated, max_keep=None, counts=None):
    """boxes (n,5) or (P,n,5) sorted by descending score -> (keep int64 (P,max_keep), num int32 (P)) on device."""
    single = boxes.dim() == 2
    b3 = boxes.unsqueeze(0) if single else boxes
    P, n, _ = b3.shape
    max_keep = n if max_keep is None else int(max_keep)
    keep = torch.zeros((P, max(max_keep, 1)), dtype=torch.int64, device=boxes.device)
    num = torch.zeros((P,), dtype=torch.int32, device=boxes.device)
    if P and n:
        cabi.call("pn2_nms_bev_f32", ptr(b3), i32(P), i32(n), i32(n), ptr(counts), f32(thresh), i32(1 if rotated else 0),
                  i32(max_keep), ptr(keep), ptr(num))
    return keep, num


def _nms(boxes, keep, thresh, rotated):
    boxes = _boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64:
        raise cabi.Pn2Error("keep must be a CPU LongTensor (iou3d_utils.py:68)")
    k, num = nms_device(boxes, thresh, rotated)
    n = int(num.item())
    keep[:n] = k[0, :n].cpu()
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, False)
