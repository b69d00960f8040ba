"""Mirror of pointrcnn/lib/utils/object3d.py: KITTI label line -> Object3d (attribute `pos`, float32;
score -1.0 when absent; object3d.py:12-36) and the label-file reader of kitti_utils.py:8-12."""
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
        self.trucation = float(label[1])
        self.occlusion = float(label[2])
        self.alpha = float(label[3])
        self.box2d = np.array([float(v) for v in label[4:8]], dtype=np.float32)
        self.h, self.w, self.l = float(label[8]), float(label[9]), float(label[10])
        self.pos = np.array([float(v) for v in label[11:14]], dtype=np.float32)
        self.dis_to_cam = np.linalg.norm(self.pos)
        self.ry = float(label[14])
        self.score = float(label[15]) if len(label) == 16 else -1.0
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

    def to_kitti_format(self):
        return '%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f' % (
            self.cls_type, self.trucation, int(self.occlusion), self.alpha, self.box2d[0], self.box2d[1], self.box2d[2],
            self.box2d[3], self.h, self.w, self.l, self.pos[0], self.pos[1], self.pos[2], self.ry)


def get_objects_from_label(label_file):
    with open(label_file, 'r') as f:
        return [Object3d(line) for line in f.readlines() if line.strip()]
