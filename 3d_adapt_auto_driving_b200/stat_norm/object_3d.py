"""Mirror of utils/object_3d.py: Object3d (KITTI label line -> attributes) and read_label.
Attribute names, dtypes (t and box2d are float32 arrays, everything else Python floats) and the
KITTI text format of to_kitti_format are the reference's (utils/object_3d.py:12-40, :116-127):
stat_norm/norm.py relies on them (float32 t mixes into float64 point arithmetic)."""
import numpy as np

_TYPE_TO_ID = {'Car': 1, 'Pedestrian': 2, 'Cyclist': 3, 'Van': 4}


def cls_type_to_id(cls_type):
    return _TYPE_TO_ID.get(cls_type, -1)


class Object3d(object):
    def __init__(self, line):
        label = line.strip().split(' ')
        self.src = line
        self.cls_type = label[0]
        self.cls_id = cls_type_to_id(self.cls_type)
        self.trucation = float(label[1])          # (sic) the reference's attribute name
        self.occlusion = float(label[2])          # 0 fully visible, 1 partly, 2 largely occluded, 3 unknown
        self.alpha = float(label[3])
        self.box2d = np.array([float(v) for v in label[4:8]], dtype=np.float32)
        self.h, self.w, self.l = float(label[8]), float(label[9]), float(label[10])
        self.t = np.array([float(v) for v in label[11:14]], dtype=np.float32)
        self.dis_to_cam = np.linalg.norm(self.t)
        self.ry = float(label[14])
        self.score = None
        if len(label) == 16:
            try:
                self.score = float(label[15])
            except ValueError:
                self.track_id = label[15]
        self.level_str = None
        self.level = self.get_obj_level()

    def get_obj_level(self):
        height = float(self.box2d[3]) - float(self.box2d[1]) + 1
        for level, name, min_h, max_trunc, max_occ in ((1, 'Easy', 40, 0.15, 0), (2, 'Moderate', 25, 0.3, 1), (3, 'Hard', 25, 0.5, 2)):
            if height >= min_h and self.trucation <= max_trunc and self.occlusion <= max_occ:
                self.level_str = name
                return level
        self.level_str = 'UnKnown'
        return 4

    def generate_corners3d(self):
        """(8, 3) corners in rect-camera coordinates, bottom face first (y = 0), then y = -h."""
        l, h, w = self.l, self.h, self.w
        x = [l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2]
        y = [0, 0, 0, 0, -h, -h, -h, -h]
        z = [w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2]
        c, s = np.cos(self.ry), np.sin(self.ry)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return np.dot(R, np.vstack([x, y, z])).T + self.t

    def to_kitti_format(self):
        vals = (self.cls_type, self.trucation, int(self.occlusion), self.alpha, self.box2d[0], self.box2d[1],
                self.box2d[2], self.box2d[3], self.h, self.w, self.l, self.t[0], self.t[1], self.t[2], self.ry)
        s = '%s %.2f %d' % vals[:3] + ''.join(' %.2f' % v for v in vals[3:])
        if self.score is not None:
            s += ' %.2f' % self.score
        return s


def read_label(label_filename):
    with open(label_filename, 'r') as f:
        return [Object3d(line) for line in f.readlines() if line.strip()]
