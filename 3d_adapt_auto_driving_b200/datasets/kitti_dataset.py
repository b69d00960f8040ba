"""Mirror of pointrcnn/lib/datasets/kitti_dataset.py: per-sample file access of a KITTI-format tree
<root>/KITTI/{ImageSets/<split>.txt, object/training/{velodyne,calib,label_2,image_2}} (:12-69)."""
import os

import numpy as np
import torch.utils.data as torch_data
from PIL import Image

from .. import calibration
from .. import object3d


class KittiDataset(torch_data.Dataset):
    def __init__(self, root_dir, split='train', subsample=-1, shuffle_subsample=None):
        self.split = split
        is_test = self.split == 'test'
        self.imageset_dir = os.path.join(root_dir, 'KITTI', 'object', 'testing' if is_test else 'training')
        split_dir = os.path.join(root_dir, 'KITTI', 'ImageSets', split + '.txt')
        self.image_idx_list = [x.strip() for x in open(split_dir).readlines()]
        if subsample > 0 and split == 'train':
            self.image_idx_list = self.image_idx_list[:subsample]
        self.num_sample = len(self.image_idx_list)
        self.image_dir = os.path.join(self.imageset_dir, 'image_2')
        self.lidar_dir = os.path.join(self.imageset_dir, 'velodyne')
        self.calib_dir = os.path.join(self.imageset_dir, 'calib')
        self.label_dir = os.path.join(self.imageset_dir, 'label_2')
        self.plane_dir = os.path.join(self.imageset_dir, 'planes')

    def get_image_shape(self, idx):
        img_file = os.path.join(self.image_dir, '%06d.png' % idx)
        assert os.path.exists(img_file)
        width, height = Image.open(img_file).size
        return height, width, 3

    def get_lidar(self, idx):
        lidar_file = os.path.join(self.lidar_dir, '%06d.bin' % idx)
        assert os.path.exists(lidar_file)
        return np.fromfile(lidar_file, dtype=np.float32).reshape(-1, 4)

    def get_calib(self, idx):
        calib_file = os.path.join(self.calib_dir, '%06d.txt' % idx)
        assert os.path.exists(calib_file)
        return calibration.Calibration(calib_file)

    def get_label(self, idx):
        label_file = os.path.join(self.label_dir, '%06d.txt' % idx)
        assert os.path.exists(label_file)
        return object3d.get_objects_from_label(label_file)

    def __len__(self):
        raise NotImplementedError

    def __getitem__(self, item):
        raise NotImplementedError
