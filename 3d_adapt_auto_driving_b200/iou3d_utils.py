"""Mirror of pointrcnn/lib/utils/iou3d/iou3d_utils.py: boxes_iou_bev, boxes_iou3d_gpu, nms_gpu,
nms_normal_gpu with the reference's signatures and return values (keep = indices into the
ORIGINAL box order, on the device).  Unlike the reference (iou3d_utils.py:68-70: CPU keep
tensor, blocking copy of the whole suppression matrix, `.cuda()` re-upload) the NMS result
never leaves the GPU."""
import torch

from . import iou3d_cuda
from . import kitti_utils


def boxes_iou_bev(boxes_a, boxes_b):
    """(M,5),(N,5) [x1,y1,x2,y2,ry] -> (M,N) rotated BEV IoU."""
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(M,7) [x,y,z,h,w,l,ry] -> (N,M) 3D IoU = BEV overlap x height overlap / union (iou3d_utils.py:21-53).
    One launch (pn2_boxes_iou3d_f32) for contiguous CUDA float32 boxes -- eval_rcnn.py calls this twice per scene in its
    recall bookkeeping; boxes_iou3d_torch is the reference's composition (one kernel + ~14 elementwise launches) and what
    tests/test_iou3d_roipool_gpu.py compares the kernel with."""
    if (boxes_a.is_cuda and boxes_b.is_cuda and boxes_a.dtype == torch.float32 and boxes_b.dtype == torch.float32
            and boxes_a.dim() == 2 and boxes_b.dim() == 2 and boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7):
        from .cabi import call, i32, ptr
        a, b = boxes_a.contiguous(), boxes_b.contiguous()
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
        call("pn2_boxes_iou3d_f32", ptr(a), i32(a.shape[0]), ptr(b), i32(b.shape[0]), ptr(out))
        return out
    return boxes_iou3d_torch(boxes_a, boxes_b)


def boxes_iou3d_torch(boxes_a, boxes_b):
    """boxes_iou3d_gpu as the reference composes it: BEV overlap kernel + torch elementwise statements."""
    boxes_a_bev = kitti_utils.boxes3d_to_bev_torch(boxes_a)
    boxes_b_bev = kitti_utils.boxes3d_to_bev_torch(boxes_b)
    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_cuda.boxes_overlap_bev_gpu(boxes_a_bev.contiguous(), boxes_b_bev.contiguous(), overlaps_bev)
    a_min = (boxes_a[:, 1] - boxes_a[:, 3]).view(-1, 1)
    a_max = boxes_a[:, 1].view(-1, 1)
    b_min = (boxes_b[:, 1] - boxes_b[:, 3]).view(1, -1)
    b_max = boxes_b[:, 1].view(1, -1)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-7)


def _nms(boxes, scores, thresh, rotated, max_keep=None):
    scores = scores.reshape(-1)      # eval_rcnn.py:626 passes (n, 1) scores; the result is .view(-1)'ed there
    order = scores.sort(0, descending=True)[1]
    sorted_boxes = boxes[order].contiguous()
    n = sorted_boxes.shape[0]
    if n == 0:
        return order
    from . import glue
    keep, num = glue.nms_raw(sorted_boxes.view(1, n, 5), None, thresh, rotated, n if max_keep is None else int(max_keep))
    return order[keep[0, :int(num.item())]].contiguous()


def nms_gpu(boxes, scores, thresh, max_keep=None):
    """rotated-IoU NMS (iou3d_utils.py:56-70). boxes (N,5), scores (N) -> kept original indices.
    max_keep (extension): stop after that many survivors."""
    return _nms(boxes, scores, thresh, True, max_keep)


def nms_normal_gpu(boxes, scores, thresh, max_keep=None):
    """axis-aligned NMS (iou3d_utils.py:73-87)."""
    return _nms(boxes, scores, thresh, False, max_keep)
