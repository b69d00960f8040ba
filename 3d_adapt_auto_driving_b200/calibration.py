"""Mirror of pointrcnn/lib/utils/calibration.py: the float32 KITTI calibration used by the PointRCNN
data path (lidar -> rectified camera -> image) and by save_kitti_format (corners3d_to_img_boxes).
Same matrix products in the same order and dtype (calibration.py:5-125)."""
import numpy as np


def get_calib_from_file(calib_file):
    with open(calib_file) as f:
        lines = f.readlines()

    def row(i):
        return np.array(lines[i].strip().split(' ')[1:], dtype=np.float32)

    return {'P2': row(2).reshape(3, 4), 'P3': row(3).reshape(3, 4), 'R0': row(4).reshape(3, 3),
            'Tr_velo2cam': row(5).reshape(3, 4)}


class Calibration(object):
    def __init__(self, calib_file):
        calib = get_calib_from_file(calib_file) if isinstance(calib_file, str) else calib_file
        self.P2, self.R0, self.V2C = calib['P2'], calib['R0'], calib['Tr_velo2cam']
        self.cu, self.cv = self.P2[0, 2], self.P2[1, 2]
        self.fu, self.fv = self.P2[0, 0], self.P2[1, 1]
        self.tx = self.P2[0, 3] / (-self.fu)
        self.ty = self.P2[1, 3] / (-self.fv)

    @staticmethod
    def cart_to_hom(pts):
        return np.hstack((pts, np.ones((pts.shape[0], 1), dtype=np.float32)))

    def lidar_to_rect(self, pts_lidar):
        return np.dot(self.cart_to_hom(pts_lidar), np.dot(self.V2C.T, self.R0.T))

    def rect_to_img(self, pts_rect):
        pts_rect_hom = self.cart_to_hom(pts_rect)
        pts_2d_hom = np.dot(pts_rect_hom, self.P2.T)
        pts_rect_hom[:, 2][pts_rect_hom[:, 2] == 0] = 1e-9
        pts_img = (pts_2d_hom[:, 0:2].T / pts_rect_hom[:, 2]).T
        pts_rect_depth = pts_2d_hom[:, 2] - self.P2.T[3, 2]
        return pts_img, pts_rect_depth

    def lidar_to_img(self, pts_lidar):
        return self.rect_to_img(self.lidar_to_rect(pts_lidar))

    def img_to_rect(self, u, v, depth_rect):
        x = ((u - self.cu) * depth_rect) / self.fu + self.tx
        y = ((v - self.cv) * depth_rect) / self.fv + self.ty
        return np.concatenate((x.reshape(-1, 1), y.reshape(-1, 1), depth_rect.reshape(-1, 1)), axis=1)

    def corners3d_to_img_boxes(self, corners3d):
        """(N,8,3) rect corners -> boxes (N,4) [x1,y1,x2,y2] and boxes_corner (N,8,2) in image coordinates."""
        sample_num = corners3d.shape[0]
        corners3d_hom = np.concatenate((corners3d, np.ones((sample_num, 8, 1))), axis=2)
        img_pts = np.matmul(corners3d_hom, self.P2.T)
        x, y = img_pts[:, :, 0] / img_pts[:, :, 2], img_pts[:, :, 1] / img_pts[:, :, 2]
        x1, y1, x2, y2 = np.min(x, axis=1), np.min(y, axis=1), np.max(x, axis=1), np.max(y, axis=1)
        boxes = np.concatenate((x1.reshape(-1, 1), y1.reshape(-1, 1), x2.reshape(-1, 1), y2.reshape(-1, 1)), axis=1)
        boxes_corner = np.concatenate((x.reshape(-1, 8, 1), y.reshape(-1, 8, 1)), axis=2)
        return boxes, boxes_corner
