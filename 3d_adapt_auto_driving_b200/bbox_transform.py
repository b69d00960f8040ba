"""Mirror of pointrcnn/lib/utils/bbox_transform.py: decode_bbox_target (:24-121) and
rotate_pc_along_y_torch (:5-21).  Bin-based box decoding; the float operation ORDER follows
the reference statement by statement so that decoded boxes are bit-identical (they feed
threshold tests and NMS)."""
import numpy as np
import torch


def rotate_pc_along_y_torch(pc, rot_angle):
    """pc (N,3+C) rotated about y by rot_angle (N): [x',z'] = [x,z] @ [[c,-s],[s,c]]^T, in place."""
    cosa = torch.cos(rot_angle).view(-1, 1)
    sina = torch.sin(rot_angle).view(-1, 1)
    R = torch.stack((torch.cat([cosa, -sina], dim=1), torch.cat([sina, cosa], dim=1)), dim=1)  # (N,2,2)
    # same batched matmul as the reference (:19-20); the columns are picked with stack / slices instead of
    # pc[:, [0, 2]] because a Python-list index uploads an index tensor from pageable memory, which
    # a CUDA-graph capture forbids
    xz = torch.stack((pc[:, 0], pc[:, 2]), dim=1).unsqueeze(dim=1)  # (N,1,2)
    out = torch.matmul(xz, R.permute(0, 2, 1)).squeeze(dim=1)
    pc[:, 0] = out[:, 0]
    pc[:, 2] = out[:, 1]
    return pc


def _bin_and_residual(pred, bin_lo, n_bins, res_lo):
    b = torch.argmax(pred[:, bin_lo:bin_lo + n_bins], dim=1)
    res = torch.gather(pred[:, res_lo:res_lo + n_bins], dim=1, index=b.unsqueeze(dim=1)).squeeze(dim=1)
    return b, res


def decode_bbox_target(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size,
                       get_xz_fine=True, get_y_by_bin=False, loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=False):
    """roi_box3d (N,3) point xyz or (N,7) roi; pred_reg (N,C) -> (N,7) [x,y,z,h,w,l,ry].
    CUDA float32 inputs go through ONE kernel (glue.decode_bbox = pn2_decode_bbox_f32, bit-identical to the torch statements:
    tests/test_glue_gpu.py) -- this is also what the unmodified eval_rcnn.py reaches through lib.utils.bbox_transform;
    anything else runs decode_bbox_target_torch, the reference's statements."""
    if (roi_box3d.is_cuda and pred_reg.is_cuda and roi_box3d.dtype == torch.float32 and pred_reg.dtype == torch.float32
            and roi_box3d.dim() == 2 and roi_box3d.shape[1] in (3, 7) and pred_reg.shape[0] > 0):
        from . import glue
        if glue.ENABLED:
            return glue.decode_bbox(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_xz_fine=get_xz_fine,
                                    get_y_by_bin=get_y_by_bin, loc_y_scope=loc_y_scope, loc_y_bin_size=loc_y_bin_size,
                                    get_ry_fine=get_ry_fine)
    return decode_bbox_target_torch(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_xz_fine,
                                    get_y_by_bin, loc_y_scope, loc_y_bin_size, get_ry_fine)


def decode_bbox_target_torch(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size,
                             get_xz_fine=True, get_y_by_bin=False, loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=False):
    """decode_bbox_target as the reference's torch statements, in the reference's operation order."""
    anchor_size = anchor_size.to(roi_box3d.device)
    nb = int(loc_scope / loc_bin_size) * 2
    nby = int(loc_y_scope / loc_y_bin_size) * 2

    x_bin = torch.argmax(pred_reg[:, 0:nb], dim=1)
    z_bin = torch.argmax(pred_reg[:, nb:nb * 2], dim=1)
    pos_x = x_bin.float() * loc_bin_size + loc_bin_size / 2 - loc_scope
    pos_z = z_bin.float() * loc_bin_size + loc_bin_size / 2 - loc_scope
    off = nb * 2
    if get_xz_fine:
        x_res = torch.gather(pred_reg[:, nb * 2:nb * 3], dim=1, index=x_bin.unsqueeze(dim=1)).squeeze(dim=1)
        z_res = torch.gather(pred_reg[:, nb * 3:nb * 4], dim=1, index=z_bin.unsqueeze(dim=1)).squeeze(dim=1)
        pos_x += x_res * loc_bin_size
        pos_z += z_res * loc_bin_size
        off = nb * 4

    if get_y_by_bin:
        y_bin, y_res = _bin_and_residual(pred_reg, off, nby, off + nby)
        pos_y = y_bin.float() * loc_y_bin_size + loc_y_bin_size / 2 - loc_y_scope + y_res * loc_y_bin_size
        pos_y = pos_y + roi_box3d[:, 1]
        off += 2 * nby
    else:
        pos_y = roi_box3d[:, 1] + pred_reg[:, off]
        off += 1

    ry_bin, ry_res_norm = _bin_and_residual(pred_reg, off, num_head_bin, off + num_head_bin)
    if get_ry_fine:
        angle_per_class = (np.pi / 2) / num_head_bin
        ry_res = ry_res_norm * (angle_per_class / 2)
        ry = (ry_bin.float() * angle_per_class + angle_per_class / 2) + ry_res - np.pi / 4
    else:
        angle_per_class = (2 * np.pi) / num_head_bin
        ry_res = ry_res_norm * (angle_per_class / 2)
        ry = (ry_bin.float() * angle_per_class + ry_res) % (2 * np.pi)
        # reference: ry[ry > np.pi] -= 2 * np.pi (:102) -- same values, but a boolean-mask index synchronises
        ry = torch.where(ry > np.pi, ry - 2 * np.pi, ry)
    off += 2 * num_head_bin

    assert off + 3 == pred_reg.shape[1]
    size_res_norm = pred_reg[:, off:off + 3]
    hwl = size_res_norm * anchor_size + anchor_size

    ret = torch.cat((pos_x.view(-1, 1), pos_y.view(-1, 1), pos_z.view(-1, 1), hwl, ry.view(-1, 1)), dim=1)
    if roi_box3d.shape[1] == 7:
        roi_ry = roi_box3d[:, 6]
        ret = rotate_pc_along_y_torch(ret, -roi_ry)
        ret[:, 6] += roi_ry
    ret[:, 0] += roi_box3d[:, 0]
    ret[:, 2] += roi_box3d[:, 2]
    return ret
