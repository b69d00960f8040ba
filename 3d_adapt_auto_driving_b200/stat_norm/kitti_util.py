"""Mirror of the part of utils/kitti_util.py that stat_norm/norm.py uses: load_velo_scan and
Calibration with the velodyne <-> reference-camera <-> rectified-camera <-> image projections
(utils/kitti_util.py:12-173).  Same matrix products in the same order (float64, np.dot), so the
projected coordinates are bit-identical to the reference's on the same BLAS."""
import numpy as np


def load_velo_scan(velo_filename):
    return np.fromfile(velo_filename, dtype=np.float32).reshape((-1, 4))


def inverse_rigid_trans(Tr):
    """inverse of a (3,4) rigid transform [R | t]  ->  [R^T | -R^T t]"""
    inv = np.zeros_like(Tr)
    inv[0:3, 0:3] = np.transpose(Tr[0:3, 0:3])
    inv[0:3, 3] = np.dot(-np.transpose(Tr[0:3, 0:3]), Tr[0:3, 3])
    return inv


class Calibration(object):
    def __init__(self, calib_filepath, from_video=False):
        if from_video:
            raise NotImplementedError("video calibration files are not on the stat_norm path")
        self.calibs = self.read_calib_file(calib_filepath)
        self.P = np.reshape(self.calibs['P2'], [3, 4])
        self.P3 = np.reshape(self.calibs['P3'], [3, 4]) if 'P3' in self.calibs else None
        self.V2C = np.reshape(self.calibs['Tr_velo_to_cam'], [3, 4])
        self.C2V = inverse_rigid_trans(self.V2C)
        self.R0 = np.reshape(self.calibs['R0_rect'], [3, 3])
        self.c_u, self.c_v = self.P[0, 2], self.P[1, 2]
        self.f_u, self.f_v = self.P[0, 0], self.P[1, 1]
        self.b_x = self.P[0, 3] / (-self.f_u)
        self.b_y = self.P[1, 3] / (-self.f_v)

    @staticmethod
    def read_calib_file(filepath):
        data = {}
        with open(filepath, 'r') as f:
            for line in f.readlines():
                line = line.rstrip()
                if not line:
                    continue
                key, value = line.split(':', 1)
                try:
                    data[key] = np.array([float(x) for x in value.split()])
                except ValueError:
                    pass
        return data

    @staticmethod
    def cart2hom(pts_3d):
        return np.hstack((pts_3d, np.ones((pts_3d.shape[0], 1))))

    def project_velo_to_ref(self, pts_3d_velo):
        return np.dot(self.cart2hom(pts_3d_velo), np.transpose(self.V2C))

    def project_ref_to_velo(self, pts_3d_ref):
        return np.dot(self.cart2hom(pts_3d_ref), np.transpose(self.C2V))

    def project_rect_to_ref(self, pts_3d_rect):
        return np.transpose(np.dot(np.linalg.inv(self.R0), np.transpose(pts_3d_rect)))

    def project_ref_to_rect(self, pts_3d_ref):
        return np.transpose(np.dot(self.R0, np.transpose(pts_3d_ref)))

    def project_rect_to_velo(self, pts_3d_rect):
        return self.project_ref_to_velo(self.project_rect_to_ref(pts_3d_rect))

    def project_velo_to_rect(self, pts_3d_velo):
        return self.project_ref_to_rect(self.project_velo_to_ref(pts_3d_velo))

    def project_rect_to_image2(self, pts_3d_rect):
        """(n,3) rect -> (n,3): image u, v and the depth term (kitti_util.py:164-173)."""
        pts_2d = np.dot(self.cart2hom(pts_3d_rect), np.transpose(self.P))
        pts_2d[:, 0] /= pts_2d[:, 2]
        pts_2d[:, 1] /= pts_2d[:, 2]
        return pts_2d

    def project_rect_to_image(self, pts_3d_rect):
        return self.project_rect_to_image2(pts_3d_rect)[:, 0:2]
