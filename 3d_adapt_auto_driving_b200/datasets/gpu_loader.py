"""GPU data path for KittiRCNNDataset (EVAL / TEST): the per-scene numpy chain of get_rpn_sample
(pointrcnn/lib/datasets/kitti_rcnn_dataset.py:249-342) -- lidar -> rect, projection into the image, the
image / PC_AREA_SCOPE filter, near / far lists and the final gather of the 16384 sampled points -- runs on
the device (csrc/scene_prepare.cu), for a whole batch of scenes per launch.  SURVEY.md 8(f) row N1.

The host keeps what defines WHICH points are sampled: the np.random draws, in the reference's order
(kitti_rcnn_dataset.py:291-320).  They depend on the cloud only through the counts (valid, near, far) the
filter kernel returns, so `draw_selection` replays exactly the calls of `_sample_indices` on index ranges
instead of index arrays (np.random.choice(n, ...) and np.random.choice(array_of_len_n, ...) consume the
generator identically) and encodes the result against the device-side lists.  With the same seed the batch
is bit-identical to collate_batch over dataset[i] (tests/test_gpu_loader_gpu.py).

    loader = GpuSceneLoader(dataset, device, batch_size=16)
    for batch in loader:                       # {'pts_input': (B,16384,3) CUDA, 'sample_id': (B,) int, 'gt_boxes3d': ...}
        ticket = detector.submit(batch['pts_input'])
"""
import concurrent.futures
import ctypes
import threading
import os

import numpy as np
import torch

from .. import cabi
from .. import kitti_utils
from ..cabi import i32, ptr
from ..config import cfg

FAR_BASE = 1 << 30


def draw_selection(n_valid, n_near, n_far, npoints, npoints_faraway, with_replace=False, rng=np.random):
    """The np.random calls of KittiRCNNDataset._sample_indices on counts -> (npoints,) int32 encoded selection
    (see scene_gather_kernel: [0, 2^30) near_list index, [2^30, 2^31) far_list index, negative = -(valid index) - 1).
    rng: the np.random module (the reference's global stream) or a np.random.RandomState seeded like it -- the same
    MT19937 algorithms, which lets per-scene-seeded draws run on several threads."""
    if npoints < n_valid:
        far_sel = np.arange(n_far, dtype=np.int64)
        if n_far > npoints_faraway:
            far_sel = rng.choice(n_far, npoints_faraway, replace=False)
        need = npoints - len(far_sel)
        if n_near < need:
            near_sel = rng.choice(n_near, need, replace=True)
        else:
            near_sel = rng.choice(n_near, need, replace=with_replace)
        choice = np.concatenate((near_sel, far_sel + FAR_BASE)) if len(far_sel) > 0 else near_sel
    else:
        choice = np.arange(n_valid, dtype=np.int64)
        if npoints > n_valid:
            missing = npoints - n_valid
            extra = rng.choice(n_valid, missing, replace=n_valid < missing)
            choice = np.concatenate((choice, extra))
        choice = -choice - 1
    order = np.arange(len(choice))
    rng.shuffle(order)                         # the same n - 1 draws as shuffling `choice` itself
    return choice[order].astype(np.int32)


def _mt_lib():
    lib = cabi.lib()
    if not getattr(lib, "_pn2_mt_ready", False):
        u32p, i32p, i64p = (ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_longlong))
        lib.pn2_mt_seed.argtypes = [ctypes.c_uint32, u32p, i32p]
        lib.pn2_mt_seed.restype = None
        lib.pn2_mt_draw_selection.argtypes = [u32p, i32p] + [ctypes.c_int] * 6 + [i32p, i32p]
        lib.pn2_mt_draw_selection.restype = ctypes.c_int
        lib._pn2_mt_ready = True
    return lib


class MTState:
    """An MT19937 state outside numpy: key (624,) uint32 + pos, as np.random.get_state() exposes it."""

    def __init__(self, key=None, pos=624):
        self.key = np.ascontiguousarray(key, np.uint32).copy() if key is not None else np.empty((624,), np.uint32)
        self.pos = ctypes.c_int32(int(pos))

    @staticmethod
    def seeded(seed):
        st = MTState()
        _mt_lib().pn2_mt_seed(ctypes.c_uint32(int(seed) & 0xffffffff), st.key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                              ctypes.byref(st.pos))
        return st

    @staticmethod
    def from_numpy_global():
        st = np.random.get_state()
        m = MTState(st[1], st[2])
        m._rest = (st[3], st[4])
        return m

    def to_numpy_global(self):
        np.random.set_state(('MT19937', self.key, int(self.pos.value)) + getattr(self, "_rest", (0, 0.0)))


def draw_selection_native(state, n_valid, n_near, n_far, npoints, npoints_faraway, with_replace, out, scratch):
    """draw_selection() by csrc/mt_select.cu on an explicit MT19937 state (GIL released during the call): the same
    selection and the same final generator state as the numpy calls.  out (npoints,) int32, scratch int32."""
    lib = _mt_lib()
    rc = lib.pn2_mt_draw_selection(state.key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(state.pos),
                                   int(n_valid), int(n_near), int(n_far), int(npoints), int(npoints_faraway),
                                   1 if with_replace else 0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   scratch.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    cabi.check(rc, "pn2_mt_draw_selection")


class GpuSceneLoader:
    """Iterates a KittiRCNNDataset in batches with the point pipeline on the GPU.  `random_select` datasets only
    (the eval_rcnn.py configuration).  Scenes are drawn in dataset order from the global np.random stream, or
    re-seeded per scene when the dataset is sharded (dataset.per_scene_seed), exactly like dataset[i]."""

    def __init__(self, dataset, device, batch_size=16, with_features=False):
        if not dataset.random_select:
            raise NotImplementedError("GpuSceneLoader mirrors the random_select path of get_rpn_sample")
        self.ds, self.device, self.batch_size = dataset, device, int(batch_size)
        self.with_features = with_features or bool(cfg.RPN.USE_INTENSITY)
        self._bufs = {}
        # host threads: file reads of the NEXT batches (they release the GIL) and, when every scene has its own seed,
        # the MT19937 draws of a batch (1.8 ms per scene for a 50 k-point permutation + the 16384 shuffle; mtrand runs
        # them without the GIL).  With the reference's single global stream the draws stay serial by definition.
        self._io = concurrent.futures.ThreadPoolExecutor(max_workers=2)
        # pinned staging ring for load_raw: one slot per job that can be in flight (3 submitted + 1 being consumed) + 1
        self._ring = {"slots": [None] * 6, "next": 0}
        self._ring_lock = threading.Lock()
        # one more thread runs prepare() itself a batch ahead, on the loader's own CUDA stream: the (serial, by the
        # reference's definition) global-stream draws then overlap the consumer's Python work instead of adding to it
        self._prep = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        self._stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None
        self._rng_pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) - 1)))

    def __len__(self):
        return (len(self.ds) + self.batch_size - 1) // self.batch_size

    def _buffers(self, b, cap):
        key = (b, cap)
        if key not in self._bufs:
            dev = self.device
            self._bufs = {key: {
                "valid": torch.empty((b, cap, 4), dtype=torch.float32, device=dev),
                "near": torch.empty((b, cap), dtype=torch.int32, device=dev),
                "far": torch.empty((b, cap), dtype=torch.int32, device=dev),
                "counts": torch.empty((b, 4), dtype=torch.int32, device=dev),
                "counts_h": torch.empty((b, 4), dtype=torch.int32).pin_memory(),
                "sel_h": torch.empty((b, self.ds.npoints), dtype=torch.int32).pin_memory(),
            }}
        return self._bufs[key]

    def _staging(self, nbytes_points, b):
        """pinned host staging (raw points, offsets, calibration), a small ring so that the H2D copies of one batch
        may still be in flight while the next batch is read; allocated once and grown on demand -- a fresh
        cudaHostAlloc per batch costs more than reading the files."""
        with self._ring_lock:            # load_raw runs on the two _io workers concurrently: slot choice must be atomic
            ring = self._ring
            k = ring["next"]
            ring["next"] = (k + 1) % len(ring["slots"])
        slot = ring["slots"][k]
        if slot is None or slot["raw"].shape[0] < nbytes_points or slot["calib"].shape[0] < b:
            pin = torch.cuda.is_available()
            cap_pts = max(nbytes_points, 1) * 5 // 4
            slot = {"raw": torch.empty((cap_pts, 4), dtype=torch.float32), "offsets": torch.empty((b + 1,), dtype=torch.int64),
                    "calib": torch.empty((b, 32), dtype=torch.float32)}
            if pin:
                slot = {k2: v.pin_memory() for k2, v in slot.items()}
            ring["slots"][k] = slot
        return slot

    def load_raw(self, indices):
        """host side of a batch: .bin clouds read straight into one pinned buffer, per-scene calibration block."""
        ds = self.ds
        ids = [int(ds.sample_id_list[index]) for index in indices]
        files = [os.path.join(ds.lidar_dir, '%06d.bin' % (sid % 200000)) for sid in ids]
        sizes = [os.path.getsize(f) // 16 for f in files]                  # (x, y, z, intensity) float32 rows
        offsets = np.zeros((len(ids) + 1,), np.int64)
        offsets[1:] = np.cumsum(sizes)
        slot = self._staging(int(offsets[-1]), len(ids))
        raw_np = slot["raw"].numpy()
        shapes = []
        calib_np = slot["calib"].numpy()
        for k, (sid, f) in enumerate(zip(ids, files)):
            with open(f, 'rb') as fh:
                fh.readinto(memoryview(raw_np[int(offsets[k]):int(offsets[k + 1])]).cast('B'))
            calib = ds.get_calib(sid)
            h, w, _ = ds.get_image_shape(sid)
            block = calib_np[k]
            block[0:12] = np.dot(calib.V2C.T, calib.R0.T).reshape(-1)      # calibration.py:55 (float32 product on the host)
            block[12:24] = calib.P2.T.reshape(-1)
            block[24], block[25] = np.float32(w), np.float32(h)
            block[26:] = 0
            shapes.append((h, w))
        slot["offsets"].numpy()[:len(ids) + 1] = offsets
        return {"raw": slot["raw"][:int(offsets[-1])], "offsets": slot["offsets"][:len(ids) + 1],
                "calib": slot["calib"][:len(ids)], "sample_id": np.array(ids, np.int32),
                "cap": int(max(sizes)), "img_shape": shapes}

    @torch.no_grad()
    def prepare(self, indices, host=None):
        """-> dict with 'pts_input' (B, npoints, 3[+1]) / 'pts_rect' / 'pts_features' on the device, 'sample_id',
        and (EVAL mode) 'gt_boxes3d' like KittiRCNNDataset.collate_batch.  host: a load_raw() result read ahead."""
        ds = self.ds
        if host is None:
            host = self.load_raw(indices)
        b, cap, npoints = len(host["sample_id"]), host["cap"], ds.npoints
        cap = (cap + 4095) // 4096 * 4096                                  # few distinct buffer shapes
        buf = self._buffers(b, cap)
        raw = host["raw"].to(self.device, non_blocking=True)
        offsets = host["offsets"].to(self.device, non_blocking=True)
        calib = host["calib"].to(self.device, non_blocking=True)
        (x0, x1), (y0, y1), (z0, z1) = [tuple(float(v) for v in r) for r in cfg.PC_AREA_SCOPE]
        d = ctypes.c_double
        cabi.call("pn2_scene_filter_f32", ptr(raw), ptr(offsets), ptr(calib), d(x0), d(x1), d(y0), d(y1), d(z0), d(z1),
                  i32(1 if cfg.PC_REDUCE_BY_RANGE else 0), ctypes.c_float(40.0), ptr(buf["valid"]), ptr(buf["near"]),
                  ptr(buf["far"]), ptr(buf["counts"]), i32(b), ctypes.c_longlong(cap), work=16.0 * raw.shape[0])
        buf["counts_h"].copy_(buf["counts"], non_blocking=True)
        torch.cuda.current_stream().synchronize()                          # the draws need the counts (loader stream only)
        counts = buf["counts_h"].numpy()
        sel = buf["sel_h"].numpy()
        def scratch_for(k):
            return np.empty((max(int(counts[k, 0]), npoints) + npoints,), np.int32)
        if ds.per_scene_seed:
            # every scene has its own generator: draw on several host threads (the native call releases the GIL)
            def draw(k):
                st = MTState.seeded((666 * 1000003 + int(host["sample_id"][k])) % (2 ** 32))
                draw_selection_native(st, counts[k, 0], counts[k, 1], counts[k, 2], npoints, ds.npoints_faraway,
                                      ds.with_replace, sel[k], scratch_for(k))
            list(self._rng_pool.map(draw, range(b)))
        else:
            # the reference's single global np.random stream, consumed in scene order: take the state out of numpy,
            # draw natively, put it back
            st = MTState.from_numpy_global()
            for k in range(b):
                draw_selection_native(st, counts[k, 0], counts[k, 1], counts[k, 2], npoints, ds.npoints_faraway,
                                      ds.with_replace, sel[k], scratch_for(k))
            st.to_numpy_global()
        sel_d = buf["sel_h"].to(self.device, non_blocking=True)
        pts = torch.empty((b, npoints, 3), dtype=torch.float32, device=self.device)
        feat = torch.empty((b, npoints), dtype=torch.float32, device=self.device) if self.with_features else None
        cabi.call("pn2_scene_gather_f32", ptr(buf["valid"]), ptr(buf["near"]), ptr(buf["far"]), ptr(sel_d), ptr(pts),
                  ptr(feat), i32(b), i32(npoints), ctypes.c_longlong(cap), work=28.0 * b * npoints)
        out = {"sample_id": host["sample_id"], "pts_rect": pts, "img_shape": host["img_shape"]}
        if feat is not None:
            out["pts_features"] = feat.unsqueeze(-1)
        out["pts_input"] = torch.cat((pts, feat.unsqueeze(-1)), dim=2) if cfg.RPN.USE_INTENSITY else pts
        if ds.mode == 'EVAL':
            gts = [kitti_utils.objs_to_boxes3d(ds.filtrate_objects(ds.get_label(int(s)))) for s in host["sample_id"]]
            max_gt = max((len(g) for g in gts), default=0)
            gt = np.zeros((b, max_gt, 7), np.float32)
            for k, g in enumerate(gts):
                gt[k, :len(g)] = g
            out["gt_boxes3d"] = gt
        return out

    def _prepare_on_stream(self, host):
        with torch.cuda.stream(self._stream):
            batch = self.prepare(None, host=host)
            batch["ready"] = torch.cuda.Event()
            batch["ready"].record(self._stream)
        return batch

    def __iter__(self):
        """Batches in dataset order.  Files are read two batches ahead and prepare() runs one batch ahead on the loader's
        stream; the consumer's current stream is made to wait for the batch before it is yielded."""
        n = len(self.ds)
        starts = list(range(0, n, self.batch_size))
        raw_ahead, prep_ahead = [], []               # load_raw futures (<= 2: the staging ring has 4 slots), prepare futures (<= 2)
        nxt = 0
        for _ in starts:
            while nxt < len(starts) and len(raw_ahead) + len(prep_ahead) < 3:
                rng_ = range(starts[nxt], min(n, starts[nxt] + self.batch_size))
                raw_ahead.append(self._io.submit(self.load_raw, rng_))
                nxt += 1
            while raw_ahead and len(prep_ahead) < 2:
                prep_ahead.append(self._prep.submit(self._prepare_on_stream, raw_ahead.pop(0).result()))
            batch = prep_ahead.pop(0).result()
            cur = torch.cuda.current_stream()
            cur.wait_event(batch.pop("ready"))
            for key in ("pts_input", "pts_rect", "pts_features"):
                if key in batch:
                    batch[key].record_stream(cur)    # allocated on the loader's stream, consumed on the caller's
            yield batch
