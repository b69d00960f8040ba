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

Fast path (default, PN2_NATIVE_DATAPATH=0 disables it): the first two steps in native code.  eval_rcnn.py owns the loop and
its DataLoader worker processes, so the data path of the unmodified script stays on the CPU -- but ten numpy passes over a
120 000-point sweep cost 10-13 ms per scene and bounded the whole script (204 scenes/s, profiles/r2_config5_n1.json).
`pn2_scene_filter_host_f32` (csrc/scene_prepare.cu) does the transform, projection and visibility test in one pass with
numpy's float32 arithmetic, and the np.random draws are replayed on the SAME MT19937 state by csrc/mt_select.cu
(np.random.get_state -> native draws -> set_state): identical samples, identical generator state afterwards
(tests/test_dataset_vs_reference_cpu.py against the live reference class, tests/test_gpu_loader_cpu.py).

Two additions the unmodified eval_rcnn.py needs: the constructor accepts `far_points` (eval_rcnn.py:862 passes it
although the reference constructor names the argument npoints_faraway -- a TypeError upstream), and scene sharding
for multi-GPU runs lives here because the script has none (PN2_SHARD_RANK / PN2_SHARD_WORLD, tools/eval_sharded.py:
the data set keeps sample_id_list[rank::world])."""
import ctypes
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
NATIVE_DATAPATH = os.environ.get("PN2_NATIVE_DATAPATH", "1") != "0"
SHARED_BATCHES = os.environ.get("PN2_SHARED_BATCHES", "1") != "0"      # worker -> main process through shared memory (_ShmArray)
_FAR_BASE = 1 << 30               # encoding of pn2_mt_draw_selection's output (datasets/gpu_loader.py)


def _native_filter(lidar, calib, img_shape):
    """visible_points() in one native pass -> (valid (k, 4) float32 [rect x, y, z, intensity], near positions, far
    positions): the rows, order and float32 values of the numpy chain"""
    from .. import cabi
    lib = cabi.lib()
    if not getattr(lib, "_pn2_filter_host_ready", False):
        f32p, i32p, f64p, i64p = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_longlong))
        lib.pn2_scene_filter_host_f32.argtypes = [f32p, ctypes.c_longlong, f32p, f32p, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                  f64p, ctypes.c_float, f32p, i32p, i32p, i64p]
        lib._pn2_filter_host_ready = True
    raw = np.ascontiguousarray(lidar, dtype=np.float32)
    n = raw.shape[0]
    m = np.ascontiguousarray(calib._velo_to_rect, dtype=np.float32)
    p = np.ascontiguousarray(calib._rect_to_image, dtype=np.float32)
    scope = np.ascontiguousarray(np.asarray(cfg.PC_AREA_SCOPE, dtype=np.float64).reshape(-1))
    valid = np.empty((max(n, 1), 4), np.float32)
    near = np.empty((max(n, 1),), np.int32)
    far = np.empty((max(n, 1),), np.int32)
    counts = np.zeros((3,), np.int64)
    fp = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    cabi.check(lib.pn2_scene_filter_host_f32(fp(raw, ctypes.c_float), n, fp(m, ctypes.c_float), fp(p, ctypes.c_float),
                                             float(img_shape[1]), float(img_shape[0]), 1 if cfg.PC_REDUCE_BY_RANGE else 0,
                                             fp(scope, ctypes.c_double), NEAR_DEPTH, fp(valid, ctypes.c_float),
                                             fp(near, ctypes.c_int32), fp(far, ctypes.c_int32), fp(counts, ctypes.c_longlong)),
               "pn2_scene_filter_host_f32")
    k, kn = int(counts[0]), int(counts[1])
    return valid[:k], near[:kn], far[:k - kn]


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


class _ShmArray(np.ndarray):
    """A batch array built inside a DataLoader WORKER.  torch's DataLoader moves torch tensors between processes through
    shared memory but pickles numpy arrays byte by byte through a pipe -- 7 MB per batch of 16 scenes, ~1.5 ms per scene of
    the main process of eval_rcnn.py, which must hand the script numpy arrays (it calls torch.from_numpy on them,
    eval_rcnn.py:498).  Pickling an instance sends its buffer as a shared-memory torch tensor instead (one copy into
    shared memory in the worker, none in the main process) and the main process receives a PLAIN ndarray on that memory."""

    def __reduce__(self):
        import torch
        return (_array_from_shared_tensor, (torch.from_numpy(np.ascontiguousarray(self).view(np.ndarray)).share_memory_(),))


def _array_from_shared_tensor(tensor):
    return tensor.numpy()


def _in_loader_worker():
    import torch.utils.data
    return torch.utils.data.get_worker_info() is not None


SHARED_BATCH_MIN_BYTES = 1 << 16     # smaller arrays (sample ids, boxes) travel by value as before


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
            stacked = np.concatenate([a[np.newaxis, ...] for a in column], axis=0)
            if stacked.nbytes >= SHARED_BATCH_MIN_BYTES and SHARED_BATCHES and _in_loader_worker():
                stacked = stacked.view(_ShmArray)
            batch[key] = stacked
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

    def _native_sample(self, sample_id):
        """visible_points + PointBudget.draw in native code: the same points in the same order, np.random left in the same
        state.  -> (xyz (npoints, 3), intensity (npoints,))"""
        from . import gpu_loader as gl
        valid, near, far = _native_filter(self.get_lidar(sample_id), self.get_calib(sample_id), self.get_image_shape(sample_id))
        if self.per_scene_seed:
            state = gl.MTState.seeded((SCENE_SEED_BASE + sample_id) % (2 ** 32))
        else:
            state = gl.MTState.from_numpy_global()
        sel = np.empty((self.npoints,), np.int32)
        scratch = np.empty((max(len(valid), self.npoints) + self.npoints,), np.int32)   # pn2_mt_draw_selection: population + selection
        gl.draw_selection_native(state, len(valid), len(near), len(far), self.npoints, self.npoints_faraway, self.with_replace,
                                 sel, scratch)
        state.to_numpy_global()
        picked = np.where(sel < 0, -sel - 1, 0).astype(np.int64)
        is_near = (sel >= 0) & (sel < _FAR_BASE)
        is_far = sel >= _FAR_BASE
        if is_near.any():
            picked[is_near] = near[sel[is_near]]
        if is_far.any():
            picked[is_far] = far[sel[is_far] - _FAR_BASE]
        chosen = valid[picked]
        return np.ascontiguousarray(chosen[:, 0:3]), np.ascontiguousarray(chosen[:, 3])

    def get_rpn_sample(self, index):
        sample_id = int(self.sample_id_list[index])
        if self.random_select and NATIVE_DATAPATH:
            xyz, intensity = self._native_sample(sample_id)
        else:
            xyz, intensity = self.visible_points(sample_id)
        if self.random_select and not NATIVE_DATAPATH:
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
