"""Mirror of pointrcnn/lib/datasets/kitti_rcnn_dataset.py, inference part: the EVAL / TEST branch of
get_rpn_sample (:249-342) -- lidar -> rectified camera, keep points that project into the image and
lie in PC_AREA_SCOPE, sample exactly `npoints` (near / far split at 40 m, far capped at
npoints_faraway, padding by duplication) with np.random in the reference's draw order -- and
collate_batch (:1125-1158).  Training branches (GT augmentation, RPN / RCNN label generation,
offline ROI sampling) are out of scope and raise.

Additions the unmodified eval_rcnn.py needs: the constructor accepts `far_points`, which
eval_rcnn.py:862 passes although the reference constructor has no such argument (a TypeError
there); it is the alias of npoints_faraway.  Scene sharding for multi-GPU runs happens here,
because eval_rcnn.py has none: with PN2_SHARD_RANK / PN2_SHARD_WORLD set (tools/eval_sharded.py)
the dataset keeps sample_id_list[rank::world]."""
import os

import numpy as np

from .kitti_dataset import KittiDataset
from ..config import cfg
from .. import kitti_utils


class KittiRCNNDataset(KittiDataset):
    def __init__(self, root_dir, npoints=16384, split='train', classes='Car', mode='TRAIN', random_select=True,
                 logger=None, rcnn_training_roi_dir=None, rcnn_training_feature_dir=None, rcnn_eval_roi_dir=None,
                 rcnn_eval_feature_dir=None, gt_database_dir=None, with_replace=False, npoints_faraway=4000,
                 subsample=-1, shuffle_subsample=False, far_points=None):
        super().__init__(root_dir=root_dir, split=split, subsample=subsample, shuffle_subsample=shuffle_subsample)
        class_sets = {'Car': ('Background', 'Car'), 'People': ('Background', 'Pedestrian', 'Cyclist'),
                      'Pedestrian': ('Background', 'Pedestrian'), 'Cyclist': ('Background', 'Cyclist')}
        assert classes in class_sets, "Invalid classes: %s" % classes
        self.classes = class_sets[classes]
        self.num_class = len(self.classes)
        self.npoints = npoints
        self.random_select = random_select
        self.logger = logger
        self.with_replace = with_replace
        self.npoints_faraway = npoints_faraway if far_points is None else far_points
        self.rcnn_eval_roi_dir = rcnn_eval_roi_dir
        self.rcnn_eval_feature_dir = rcnn_eval_feature_dir
        assert mode in ['TRAIN', 'EVAL', 'TEST'], 'Invalid mode: %s' % mode
        self.mode = mode
        if mode == 'TRAIN':
            raise NotImplementedError("training data paths are out of scope of the inference package")
        if not cfg.RPN.ENABLED:
            raise NotImplementedError("offline RCNN evaluation from saved proposals is not on the eval_rcnn.py rcnn path")
        self.sample_id_list = [int(sample_id) for sample_id in self.image_idx_list]
        rank, world = int(os.environ.get("PN2_SHARD_RANK", "0")), int(os.environ.get("PN2_SHARD_WORLD", "1"))
        if world > 1:
            self.sample_id_list = self.sample_id_list[rank::world]
        # The reference draws every scene's subsampling from ONE np.random stream in scene order
        # (eval_rcnn.py:467 seeds it once), so which points a scene keeps depends on all scenes before
        # it -- impossible to reproduce on a shard without loading every other shard's scenes.  Sharded
        # runs (and PN2_PER_SCENE_SEED=1) therefore re-seed per scene: results are then identical for
        # every world size, and differ from the single-stream order only in which points are sampled.
        self.per_scene_seed = world > 1 or os.environ.get("PN2_PER_SCENE_SEED", "0") == "1"
        if self.logger is not None:
            self.logger.info('Load testing samples from %s' % self.imageset_dir)
            self.logger.info('Done: total test samples %d' % len(self.sample_id_list))

    def get_image_shape(self, idx):
        return super().get_image_shape(idx % 200000)

    def get_calib(self, idx):
        return super().get_calib(idx % 200000)

    def get_label(self, idx):
        return super().get_label(idx % 200000)

    def filtrate_objects(self, obj_list):
        return [obj for obj in obj_list if obj.cls_type in self.classes]

    @staticmethod
    def get_valid_flag(pts_rect, pts_img, pts_rect_depth, img_shape):
        """in the image and (PC_REDUCE_BY_RANGE) inside PC_AREA_SCOPE (:201-222)"""
        val_flag_1 = np.logical_and(pts_img[:, 0] >= 0, pts_img[:, 0] < img_shape[1])
        val_flag_2 = np.logical_and(pts_img[:, 1] >= 0, pts_img[:, 1] < img_shape[0])
        pts_valid_flag = np.logical_and(np.logical_and(val_flag_1, val_flag_2), pts_rect_depth >= 0)
        if cfg.PC_REDUCE_BY_RANGE:
            x_range, y_range, z_range = cfg.PC_AREA_SCOPE
            pts_x, pts_y, pts_z = pts_rect[:, 0], pts_rect[:, 1], pts_rect[:, 2]
            range_flag = (pts_x >= x_range[0]) & (pts_x <= x_range[1]) & (pts_y >= y_range[0]) & (pts_y <= y_range[1]) \
                & (pts_z >= z_range[0]) & (pts_z <= z_range[1])
            pts_valid_flag = pts_valid_flag & range_flag
        return pts_valid_flag

    def __len__(self):
        return len(self.sample_id_list)

    def __getitem__(self, index):
        return self.get_rpn_sample(index)

    def _sample_indices(self, pts_rect):
        """exactly self.npoints indices; the np.random draws are the reference's, in its order (:291-320)"""
        if self.npoints < len(pts_rect):
            pts_near_flag = pts_rect[:, 2] < 40.0
            far_idxs_choice = np.where(pts_near_flag == 0)[0]
            if len(far_idxs_choice) > self.npoints_faraway:
                far_idxs_choice = np.random.choice(far_idxs_choice, self.npoints_faraway, replace=False)
            near_idxs = np.where(pts_near_flag == 1)[0]
            need = self.npoints - len(far_idxs_choice)
            if len(near_idxs) < need:
                near_idxs_choice = np.random.choice(near_idxs, need, replace=True)
            else:
                near_idxs_choice = np.random.choice(near_idxs, need, replace=self.with_replace)
            choice = np.concatenate((near_idxs_choice, far_idxs_choice), axis=0) if len(far_idxs_choice) > 0 \
                else near_idxs_choice
            np.random.shuffle(choice)
        else:
            choice = np.arange(0, len(pts_rect), dtype=np.int32)
            if self.npoints > len(pts_rect):
                missing = self.npoints - len(pts_rect)
                extra_choice = np.random.choice(choice, missing, replace=len(choice) < missing)
                choice = np.concatenate((choice, extra_choice), axis=0)
            np.random.shuffle(choice)
        return choice

    def get_rpn_sample(self, index):
        sample_id = int(self.sample_id_list[index])
        calib = self.get_calib(sample_id)
        img_shape = self.get_image_shape(sample_id)
        pts_lidar = self.get_lidar(sample_id)
        pts_rect = calib.lidar_to_rect(pts_lidar[:, 0:3])
        pts_intensity = pts_lidar[:, 3]
        pts_img, pts_rect_depth = calib.rect_to_img(pts_rect)
        pts_valid_flag = self.get_valid_flag(pts_rect, pts_img, pts_rect_depth, img_shape)
        pts_rect = pts_rect[pts_valid_flag][:, 0:3]
        pts_intensity = pts_intensity[pts_valid_flag]
        if self.random_select:
            if self.per_scene_seed:
                np.random.seed((666 * 1000003 + sample_id) % (2 ** 32))
            choice = self._sample_indices(pts_rect)
            ret_pts_rect = pts_rect[choice, :]
            ret_pts_intensity = pts_intensity[choice] - 0.5          # intensity to [-0.5, 0.5]
        else:
            ret_pts_rect = pts_rect
            ret_pts_intensity = pts_intensity - 0.5
        ret_pts_features = ret_pts_intensity.reshape(-1, 1)
        sample_info = {'sample_id': sample_id, 'random_select': self.random_select}
        pts_input = np.concatenate((ret_pts_rect, ret_pts_features), axis=1) if cfg.RPN.USE_INTENSITY else ret_pts_rect
        sample_info['pts_input'] = pts_input
        sample_info['pts_rect'] = ret_pts_rect
        sample_info['pts_features'] = ret_pts_features
        if self.mode == 'TEST':
            return sample_info
        if not cfg.RPN.FIXED:
            raise NotImplementedError("RPN training labels are out of scope of the inference package")
        gt_obj_list = self.filtrate_objects(self.get_label(sample_id))
        sample_info['gt_boxes3d'] = kitti_utils.objs_to_boxes3d(gt_obj_list)
        return sample_info

    def collate_batch(self, batch):
        batch_size = len(batch)
        ans_dict = {}
        for key in batch[0].keys():
            if key in ('gt_boxes3d', 'roi_boxes3d'):
                max_gt = max(len(batch[k][key]) for k in range(batch_size))
                batch_gt_boxes3d = np.zeros((batch_size, max_gt, 7), dtype=np.float32)
                for i in range(batch_size):
                    batch_gt_boxes3d[i, :len(batch[i][key]), :] = batch[i][key]
                ans_dict[key] = batch_gt_boxes3d
                continue
            if isinstance(batch[0][key], np.ndarray):
                ans_dict[key] = np.concatenate([batch[k][key][np.newaxis, ...] for k in range(batch_size)], axis=0)
            else:
                ans_dict[key] = [batch[k][key] for k in range(batch_size)]
                if isinstance(batch[0][key], int):
                    ans_dict[key] = np.array(ans_dict[key], dtype=np.int32)
                elif isinstance(batch[0][key], float):
                    ans_dict[key] = np.array(ans_dict[key], dtype=np.float32)
        return ans_dict
