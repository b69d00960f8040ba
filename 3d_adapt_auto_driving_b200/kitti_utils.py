"""The torch helpers of pointrcnn/lib/utils/kitti_utils.py that sit on the inference path:
boxes3d_to_bev_torch (:134-147), enlarge_box3d (:150-160), rotate_pc_along_y_torch (:45-63),
boxes3d_to_corners3d_torch / boxes3d_to_corners3d (numpy, used by eval_rcnn.py:78)."""
import numpy as np
import torch


def boxes3d_to_bev_torch(boxes3d):
    """(N,7) [x,y,z,h,w,l,ry] -> (N,5) [x1,y1,x2,y2,ry] in the x-z plane."""
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    half_l, half_w = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    return torch.stack((cu - half_l, cv - half_w, cu + half_l, cv + half_w, boxes3d[:, 6]), dim=1)


def enlarge_box3d(boxes3d, extra_width):
    large = boxes3d.copy() if isinstance(boxes3d, np.ndarray) else boxes3d.clone()
    large[:, 3:6] += extra_width * 2
    large[:, 1] += extra_width
    return large


def rotate_pc_along_y_torch(pc, rot_angle):
    """pc (N,S,3+C) rotated about y by rot_angle (N), in place on columns 0 and 2 -- the batched
    matmul of the reference (kitti_utils.py:45-63): [x', z'] = [x, z] @ [[cos, -sin], [sin, cos]]^T.
    The reference calls it once per scene in a Python loop (rcnn_net.py:150-152); rows are
    independent, so one call over all B*M ROIs gives the same values."""
    cosa = torch.cos(rot_angle).view(-1, 1)
    sina = torch.sin(rot_angle).view(-1, 1)
    R = torch.cat((torch.cat([cosa, -sina], dim=1).unsqueeze(dim=1), torch.cat([sina, cosa], dim=1).unsqueeze(dim=1)), dim=1)
    xz = torch.stack((pc[:, :, 0], pc[:, :, 2]), dim=2)                   # (N,S,2)
    out = torch.matmul(xz, R.permute(0, 2, 1))
    pc[:, :, 0] = out[:, :, 0]
    pc[:, :, 2] = out[:, :, 1]
    return pc


# corner order of the KITTI convention (kitti_utils.py:72-77): bottom face first (y = 0), then
# the top face (y = -h); signs of (l/2, w/2) per corner
_CORNER_SX = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float32)
_CORNER_SZ = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float32)
_CORNER_TOP = np.array([0, 0, 0, 0, 1, 1, 1, 1], np.float32)


def boxes3d_to_corners3d(boxes3d, rotate=True):
    """numpy (N,7) [x,y,z,h,w,l,ry] -> (N,8,3) corners in rect-camera coordinates
    (kitti_utils.py:66-98; y points down, the box origin is the bottom-face centre).
    Bit-identical to the reference, whose rotation is a batched np.matmul of the float32 (N,8,3) offsets with a
    TRANSPOSED VIEW of a (3,3,N) rotation stack: the corners end up in the %.4f text of eval_rcnn.py:76-101 through
    the image projection, so the same numpy routine is applied to operands of the same dtype and memory layout
    (one rounding per product-sum differs between matmul's accumulation and x*c + z*s written out)."""
    b = np.asarray(boxes3d)
    n = b.shape[0]
    h, w, l = b[:, 3:4], b[:, 4:5], b[:, 5:6]
    off = np.empty((n, 8, 3), np.float32)                    # (x, y, z) offsets of the corners before rotation
    off[:, :, 0] = (l / 2.0) * _CORNER_SX
    off[:, :, 1] = (-h) * _CORNER_TOP
    off[:, :, 2] = (w / 2.0) * _CORNER_SZ
    if rotate:
        ry = b[:, 6]
        c, s = np.cos(ry), np.sin(ry)
        rot = np.zeros((3, 3, ry.size), np.result_type(c.dtype, np.float32))
        rot[0, 0], rot[0, 2] = c, -s
        rot[1, 1] = 1
        rot[2, 0], rot[2, 2] = s, c
        off = np.matmul(off, rot.transpose(2, 0, 1))         # (N,8,3) @ (N,3,3): rows [x y z] times R
    out = np.stack((b[:, 0:1] + off[:, :, 0], b[:, 1:2] + off[:, :, 1], b[:, 2:3] + off[:, :, 2]), axis=2)
    return out.astype(np.float32)


def objs_to_boxes3d(obj_list):
    """[Object3d] -> (N,7) float32 [x, y, z, h, w, l, ry] (kitti_utils.py:180-185)"""
    boxes3d = np.zeros((len(obj_list), 7), dtype=np.float32)
    for k, obj in enumerate(obj_list):
        boxes3d[k, 0:3], boxes3d[k, 3], boxes3d[k, 4], boxes3d[k, 5], boxes3d[k, 6] = obj.pos, obj.h, obj.w, obj.l, obj.ry
    return boxes3d


def get_objects_from_label(label_file):
    from . import object3d
    return object3d.get_objects_from_label(label_file)
