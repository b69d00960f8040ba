"""Inference-side data set with the interface of pointrcnn/lib/datasets/kitti_rcnn_dataset.py (KittiRCNNDataset):
what eval_rcnn.py:851-866 constructs and iterates.  Only the EVAL / TEST branch of get_rpn_sample (:249-342) and
collate_batch (:1125-1158) exist here; everything that serves training (GT augmentation, label generation, offline
ROI sampling) raises.

The file is organised around three independent steps instead of the reference's single method:

    visible_points()      lidar -> rectified camera, image / PC_AREA_SCOPE visibility (get_valid_flag, :201-222)
    PointBudget.draw()    which `npoints` of the visible points a scene keeps (:291-320).  The ONLY contract with the
                          reference here is the sequence of np.random calls (choice / choice / shuffle and their
                          arguments): scenes are sampled from one global stream, so every draw must consume it
                          identically (tests/test_dataset_vs_reference_cpu.py compares samples with the live class).
    stack_samples()       list of per-scene dicts -> batch dict (collate_batch)

Two additions the unmodified eval_rcnn.py needs: the constructor accepts `far_points` (eval_rcnn.py:862 passes it
although the reference constructor names the argument npoints_faraway -- a TypeError upstream), and scene sharding
for multi-GPU runs lives here because the script has none (PN2_SHARD_RANK / PN2_SHARD_WORLD, tools/eval_sharded.py:
the data set keeps sample_id_list[rank::world])."""
import os

import numpy as np

from .kitti_dataset import KittiDataset
from ..config import cfg
from .. import kitti_utils

CLASS_GROUPS = {
    'Car': ('Background', 'Car'),
    'People': ('Background', 'Pedestrian', 'Cyclist'),
    'Pedestrian': ('Background', 'Pedestrian'),
    'Cyclist': ('Background', 'Cyclist'),
}
NEAR_DEPTH = 40.0                 # metres: the near / far split of the point budget (:293)
BOX_KEYS = ('gt_boxes3d', 'roi_boxes3d')
SCENE_SEED_BASE = 666 * 1000003   # per-scene seeds of sharded runs: (base + sample_id) mod 2^32


def _inside(values, bounds):
    return (values >= bounds[0]) & (values <= bounds[1])


class PointBudget:
    """Reduce / pad a scene to exactly `npoints` points the way :291-320 does: at most `far_cap` points beyond 40 m,
    the rest from the near ones (with replacement only when there are too few), padding by duplication when the
    scene is smaller than the budget, and a final shuffle.  Index arithmetic is ours; the np.random calls, their
    order and their arguments are the reference's."""

    def __init__(self, npoints, far_cap, with_replace):
        self.npoints, self.far_cap, self.with_replace = npoints, far_cap, with_replace

    def _reduce(self, depth, rng):
        near_mask = depth < NEAR_DEPTH
        far = np.flatnonzero(near_mask == 0)
        near = np.flatnonzero(near_mask == 1)
        if far.size > self.far_cap:
            far = rng.choice(far, self.far_cap, replace=False)
        want_near = self.npoints - far.size
        near = rng.choice(near, want_near, replace=True if near.size < want_near else self.with_replace)
        return np.concatenate((near, far), axis=0) if far.size > 0 else near

    def _pad(self, count, rng):
        keep = np.arange(0, count, dtype=np.int32)
        short = self.npoints - count
        if short > 0:
            keep = np.concatenate((keep, rng.choice(keep, short, replace=count < short)), axis=0)
        return keep

    def draw(self, pts_rect, rng=np.random):
        count = len(pts_rect)
        picked = self._reduce(pts_rect[:, 2], rng) if self.npoints < count else self._pad(count, rng)
        rng.shuffle(picked)
        return picked


def stack_samples(samples):
    """collate_batch (:1125-1158): box lists are zero-padded to the longest of the batch, arrays are stacked, Python
    ints / floats become int32 / float32 vectors, anything else stays a list."""
    first = samples[0]
    batch = {}
    for key, probe in first.items():
        column = [s[key] for s in samples]
        if key in BOX_KEYS:
            padded = np.zeros((len(samples), max(len(boxes) for boxes in column), 7), dtype=np.float32)
            for row, boxes in zip(padded, column):
                row[:len(boxes)] = boxes
            batch[key] = padded
        elif isinstance(probe, np.ndarray):
            batch[key] = np.concatenate([a[np.newaxis, ...] for a in column], axis=0)
        elif isinstance(probe, int):
            batch[key] = np.array(column, dtype=np.int32)
        elif isinstance(probe, float):
            batch[key] = np.array(column, dtype=np.float32)
        else:
            batch[key] = column
    return batch


class KittiRCNNDataset(KittiDataset):
    def __init__(self, root_dir, npoints=16384, split='train', classes='Car', mode='TRAIN', random_select=True,
                 logger=None, rcnn_training_roi_dir=None, rcnn_training_feature_dir=None, rcnn_eval_roi_dir=None,
                 rcnn_eval_feature_dir=None, gt_database_dir=None, with_replace=False, npoints_faraway=4000,
                 subsample=-1, shuffle_subsample=False, far_points=None):
        super().__init__(root_dir=root_dir, split=split, subsample=subsample, shuffle_subsample=shuffle_subsample)
        if classes not in CLASS_GROUPS:
            raise AssertionError("Invalid classes: %s" % classes)
        if mode not in ('TRAIN', 'EVAL', 'TEST'):
            raise AssertionError('Invalid mode: %s' % mode)
        if mode == 'TRAIN':
            raise NotImplementedError("training data paths are out of scope of the inference package")
        if not cfg.RPN.ENABLED:
            raise NotImplementedError("offline RCNN evaluation from saved proposals is not on the eval_rcnn.py rcnn path")
        self.mode, self.logger = mode, logger
        self.classes = CLASS_GROUPS[classes]
        self.num_class = len(self.classes)
        self.npoints, self.random_select, self.with_replace = npoints, random_select, with_replace
        self.npoints_faraway = far_points if far_points is not None else npoints_faraway
        self.rcnn_eval_roi_dir, self.rcnn_eval_feature_dir = rcnn_eval_roi_dir, rcnn_eval_feature_dir

        ids = [int(name) for name in self.image_idx_list]
        rank = int(os.environ.get("PN2_SHARD_RANK", "0"))
        world = int(os.environ.get("PN2_SHARD_WORLD", "1"))
        self.sample_id_list = ids[rank::world] if world > 1 else ids
        # The reference samples every scene from ONE np.random stream in scene order (eval_rcnn.py:467 seeds it once),
        # so a scene's points depend on all scenes before it -- not reproducible on a shard.  Sharded runs (and
        # PN2_PER_SCENE_SEED=1) re-seed per scene instead: identical results for every world size, different from the
        # single-stream run only in WHICH points are sampled (a documented deviation).
        self.per_scene_seed = world > 1 or os.environ.get("PN2_PER_SCENE_SEED", "0") == "1"
        if logger is not None:
            logger.info('Load testing samples from %s' % self.imageset_dir)
            logger.info('Done: total test samples %d' % len(self.sample_id_list))

    # ---- file access: ids of augmented scenes wrap onto their source scene (:80-93) ----
    def get_image_shape(self, idx):
        return super().get_image_shape(idx % 200000)

    def get_calib(self, idx):
        return super().get_calib(idx % 200000)

    def get_label(self, idx):
        return super().get_label(idx % 200000)

    def filtrate_objects(self, obj_list):
        return [obj for obj in obj_list if obj.cls_type in self.classes]

    # ---- visibility ----
    @staticmethod
    def get_valid_flag(pts_rect, pts_img, pts_rect_depth, img_shape):
        """points that project into the image with non-negative depth and (cfg.PC_REDUCE_BY_RANGE) lie inside
        cfg.PC_AREA_SCOPE (:201-222)"""
        height, width = img_shape[0], img_shape[1]
        u, v = pts_img[:, 0], pts_img[:, 1]
        keep = (u >= 0) & (u < width) & (v >= 0) & (v < height) & (pts_rect_depth >= 0)
        if cfg.PC_REDUCE_BY_RANGE:
            for axis, bounds in enumerate(cfg.PC_AREA_SCOPE):
                keep &= _inside(pts_rect[:, axis], bounds)
        return keep

    def visible_points(self, sample_id):
        """-> (rect xyz (n, 3), intensity (n,)) of the scene's visible points"""
        calib = self.get_calib(sample_id)
        lidar = self.get_lidar(sample_id)
        rect = calib.lidar_to_rect(lidar[:, 0:3])
        img, depth = calib.rect_to_img(rect)
        keep = self.get_valid_flag(rect, img, depth, self.get_image_shape(sample_id))
        return rect[keep][:, 0:3], lidar[:, 3][keep]

    def _sample_indices(self, pts_rect):
        """exactly self.npoints indices into pts_rect, drawn from the global np.random stream (PointBudget)"""
        return PointBudget(self.npoints, self.npoints_faraway, self.with_replace).draw(pts_rect)

    # ---- samples ----
    def __len__(self):
        return len(self.sample_id_list)

    def __getitem__(self, index):
        return self.get_rpn_sample(index)

    def get_rpn_sample(self, index):
        sample_id = int(self.sample_id_list[index])
        xyz, intensity = self.visible_points(sample_id)
        if self.random_select:
            if self.per_scene_seed:
                np.random.seed((SCENE_SEED_BASE + sample_id) % (2 ** 32))
            picked = self._sample_indices(xyz)
            xyz, intensity = xyz[picked, :], intensity[picked]
        features = (intensity - 0.5).reshape(-1, 1)                  # intensity to [-0.5, 0.5] (:324)
        sample = {
            'sample_id': sample_id,
            'random_select': self.random_select,
            'pts_input': np.concatenate((xyz, features), axis=1) if cfg.RPN.USE_INTENSITY else xyz,
            'pts_rect': xyz,
            'pts_features': features,
        }
        if self.mode == 'TEST':
            return sample
        if not cfg.RPN.FIXED:
            raise NotImplementedError("RPN training labels are out of scope of the inference package")
        sample['gt_boxes3d'] = kitti_utils.objs_to_boxes3d(self.filtrate_objects(self.get_label(sample_id)))
        return sample

    def collate_batch(self, batch):
        return stack_samples(batch)
