"""Drop-in for evaluate/rotate_iou.py: `rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0)`
with the reference's contract (rotate_iou.py:294-329): any float dtype in, cast to float32 for the
kernel, (N, K) result returned as FLOAT32 for every input dtype -- `box_dtype` is captured (:311) but the final
`iou.astype(boxes.dtype)` (:329) runs after `boxes` was rebound to its float32 cast (:312); an empty N or K returns
float32 zeros early (:315-316).  The numba JIT kernel is replaced by the sm_100a kernel behind pn2_rotate_iou_eval_f32
(csrc/rotate_iou.cu), bit-exact with it; host <-> device copies go through torch on the selected
device.  Importers do `from rotate_iou import rotate_iou_gpu_eval` (evaluate/eval2.py:4)."""
import numpy as np
import torch

from . import cabi
from .cabi import i32, ptr


def div_up(m, n):
    return m // n + (m % n > 0)


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """boxes (N,5), query_boxes (K,5): [cx, cy, w, h, angle (clockwise positive)] -> (N,K)."""
    boxes = np.asarray(boxes)
    query_boxes = np.asarray(query_boxes)
    b32 = np.ascontiguousarray(boxes.astype(np.float32)).reshape(-1, 5) if boxes.size else boxes.astype(np.float32)
    q32 = np.ascontiguousarray(query_boxes.astype(np.float32)).reshape(-1, 5) if query_boxes.size else query_boxes.astype(np.float32)
    N, K = boxes.shape[0], query_boxes.shape[0]
    iou = np.zeros((N, K), dtype=np.float32)
    if N == 0 or K == 0:
        return iou
    if not torch.cuda.is_available():
        raise cabi.Pn2Error("rotate_iou_gpu_eval needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", int(device_id))
    with torch.cuda.device(dev):
        db = torch.from_numpy(b32).to(dev)
        dq = torch.from_numpy(q32).to(dev)
        out = torch.empty((N, K), dtype=torch.float32, device=dev)
        cabi.call("pn2_rotate_iou_eval_f32", ptr(db), i32(N), ptr(dq), i32(K), ptr(out), i32(criterion),
                  work=float(N) * K)
        iou = out.cpu().numpy()
    return iou.astype(np.float32, copy=False)
