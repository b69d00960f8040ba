"""File access of a KITTI-format tree, with the interface of pointrcnn/lib/datasets/kitti_dataset.py:12-69:

    <root>/KITTI/ImageSets/<split>.txt                      one six-digit sample id per line
    <root>/KITTI/object/{training|testing}/velodyne/######.bin   float32 (x, y, z, intensity)
                                          /calib/######.txt, /label_2/######.txt, /image_2/######.png, /planes/

Attribute and method names are the ones KittiRCNNDataset and eval_rcnn.py use."""
import os

import numpy as np
import torch.utils.data as torch_data
from PIL import Image

from .. import calibration
from .. import object3d

_SUBDIRS = {'image_dir': 'image_2', 'lidar_dir': 'velodyne', 'calib_dir': 'calib', 'label_dir': 'label_2',
            'plane_dir': 'planes'}


class KittiDataset(torch_data.Dataset):
    def __init__(self, root_dir, split='train', subsample=-1, shuffle_subsample=None):
        self.split = split
        kitti = os.path.join(root_dir, 'KITTI')
        self.imageset_dir = os.path.join(kitti, 'object', 'testing' if split == 'test' else 'training')
        for attr, sub in _SUBDIRS.items():
            setattr(self, attr, os.path.join(self.imageset_dir, sub))
        with open(os.path.join(kitti, 'ImageSets', split + '.txt')) as f:
            ids = [line.strip() for line in f.readlines()]
        if split == 'train' and subsample > 0:
            ids = ids[:subsample]
        self.image_idx_list = ids
        self.num_sample = len(ids)

    @staticmethod
    def _existing(directory, idx, ext):
        path = os.path.join(directory, '%06d%s' % (idx, ext))
        assert os.path.exists(path), path
        return path

    def get_image_shape(self, idx):
        """(height, width, 3) of image_2/######.png.  Only the size is needed (kitti_dataset.py:50-55 opens the image for
        it): a PNG stores it in the 8 bytes after the 16-byte signature + IHDR header, which is a 24-byte read instead of a
        PIL open per scene in the main loop of eval_rcnn.py; anything that is not a PNG goes through PIL."""
        path = self._existing(self.image_dir, idx, '.png')
        with open(path, 'rb') as f:
            head = f.read(24)
        if len(head) == 24 and head[:8] == b'\x89PNG\r\n\x1a\n' and head[12:16] == b'IHDR':
            return int.from_bytes(head[20:24], 'big'), int.from_bytes(head[16:20], 'big'), 3
        with Image.open(path) as img:
            width, height = img.size
        return height, width, 3

    def get_lidar(self, idx):
        return np.fromfile(self._existing(self.lidar_dir, idx, '.bin'), dtype=np.float32).reshape(-1, 4)

    def get_calib(self, idx):
        return calibration.Calibration(self._existing(self.calib_dir, idx, '.txt'))

    def get_label(self, idx):
        return object3d.get_objects_from_label(self._existing(self.label_dir, idx, '.txt'))

    def __len__(self):
        raise NotImplementedError

    def __getitem__(self, item):
        raise NotImplementedError
