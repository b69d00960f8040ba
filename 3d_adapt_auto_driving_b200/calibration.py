"""KITTI calibration in float32 for the PointRCNN data path (velodyne -> rectified camera -> image plane) and for the
result writer (3-D box corners -> 2-D boxes).  Interface of pointrcnn/lib/utils/calibration.py:5-125: the attribute
names (P2, R0, V2C, cu, cv, fu, fv, tx, ty) and method names are what the dataset classes and eval_rcnn.py use.

Every projection is "append a homogeneous 1, multiply by a (4, 3) matrix"; the two matrices are fixed per scene, so
they are composed once here (`_velo_to_rect`, `_rect_to_image`) and the methods share one helper.  Results are
bit-identical to the reference's per-call products -- same operands, same dtype, same memory layout handed to BLAS
(`P2.T` stays a transposed view on purpose) -- which tests/test_host_utils_cpu.py and
tests/test_dataset_vs_reference_cpu.py check against the reference module itself."""
import numpy as np

_ROWS = {'P2': (2, (3, 4)), 'P3': (3, (3, 4)), 'R0': (4, (3, 3)), 'Tr_velo2cam': (5, (3, 4))}   # line index, shape


def get_calib_from_file(calib_file):
    """calib/######.txt -> {'P2', 'P3', 'R0', 'Tr_velo2cam'} float32; rows are addressed by position, as KITTI fixes it"""
    with open(calib_file) as f:
        lines = f.readlines()
    return {key: np.array(lines[idx].strip().split(' ')[1:], dtype=np.float32).reshape(shape)
            for key, (idx, shape) in _ROWS.items()}


def _append_one(pts):
    """(N, k) -> (N, k + 1) with a float32 column of ones (numpy promotes float64 inputs, like the reference)"""
    return np.hstack((pts, np.ones((pts.shape[0], 1), dtype=np.float32)))


class Calibration(object):
    def __init__(self, calib_file):
        calib = get_calib_from_file(calib_file) if isinstance(calib_file, str) else calib_file
        self.P2, self.R0, self.V2C = calib['P2'], calib['R0'], calib['Tr_velo2cam']
        (self.fu, _, self.cu, bx), (_, self.fv, self.cv, by) = self.P2[0], self.P2[1]
        self.tx, self.ty = bx / (-self.fu), by / (-self.fv)
        self._velo_to_rect = np.dot(self.V2C.T, self.R0.T)        # (4, 3): [x y z 1] @ . = rectified camera xyz
        self._rect_to_image = self.P2.T                          # (4, 3) view: [x y z 1] @ . = (u * d, v * d, d + P2[2, 3])

    cart_to_hom = staticmethod(_append_one)

    def lidar_to_rect(self, pts_lidar):
        return np.dot(_append_one(pts_lidar), self._velo_to_rect)

    def rect_to_img(self, pts_rect):
        """-> pixel coordinates (N, 2) and depth in the rectified camera frame (N,); a zero z divides by 1e-9"""
        hom = _append_one(pts_rect)
        projected = np.dot(hom, self._rect_to_image)
        z = hom[:, 2]
        z[z == 0] = 1e-9
        return (projected[:, 0:2].T / z).T, projected[:, 2] - self._rect_to_image[3, 2]

    def lidar_to_img(self, pts_lidar):
        return self.rect_to_img(self.lidar_to_rect(pts_lidar))

    def img_to_rect(self, u, v, depth_rect):
        x = ((u - self.cu) * depth_rect) / self.fu + self.tx
        y = ((v - self.cv) * depth_rect) / self.fv + self.ty
        return np.stack((x.reshape(-1), y.reshape(-1), depth_rect.reshape(-1)), axis=1)

    def corners3d_to_img_boxes(self, corners3d):
        """(N, 8, 3) box corners -> 2-D boxes (N, 4) [x1, y1, x2, y2] and the projected corners (N, 8, 2)"""
        hom = np.concatenate((corners3d, np.ones((corners3d.shape[0], 8, 1))), axis=2)           # float64 ones: promotes
        uvd = np.matmul(hom, self._rect_to_image)
        uv = np.stack((uvd[:, :, 0] / uvd[:, :, 2], uvd[:, :, 1] / uvd[:, :, 2]), axis=2)       # (N, 8, 2)
        boxes = np.concatenate((uv.min(axis=1), uv.max(axis=1)), axis=1)
        return boxes, uv
