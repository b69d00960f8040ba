"""Scene-sharded multi-GPU evaluation: one process per GPU, no data-path collective, ONE all_gather of
the detection records at the end (BASELINE.json north_star).

The reference has no multi-GPU inference (nn.DataParallel in training only, train_rcnn.py:207).
Scenes are independent, so rank r of W evaluates sample_id_list[r::W] (the dataset reads
PN2_SHARD_RANK / PN2_SHARD_WORLD, datasets/kitti_rcnn_dataset.py) with the unmodified eval_rcnn.py
writing into a per-rank output directory -- its "dump empty files" loop (eval_rcnn.py:638-649) would
otherwise create empty results for scenes a rank does not own.  Afterwards every rank packs the
KITTI result lines of its own scenes into a fixed-width float64 tensor, a single all_gather
(NCCL over NVLink on GPUs, gloo in the CPU tests) moves them, and rank 0 writes the merged
result directory.  Payload: scenes x 100 x 13 doubles, a few MB for the 7481-scene KITTI val set."""
import os

import numpy as np
import torch
import torch.distributed as dist

FIELDS = 13          # alpha, x1, y1, x2, y2, h, w, l, x, y, z, ry, score  (eval_rcnn.py:96-100)
MAX_DET = 100        # RPN_POST_NMS_TOP_N rois per scene bounds the detections per scene


def shard_ids(sample_ids, rank, world):
    return list(sample_ids)[rank::world]


def pack_result_dir(final_dir, sample_ids, max_det=MAX_DET):
    """KITTI txt files of `sample_ids` -> (records (n, max_det, FIELDS) float64, counts (n,) int64)."""
    rec = np.zeros((len(sample_ids), max_det, FIELDS), np.float64)
    cnt = np.zeros((len(sample_ids),), np.int64)
    for i, sid in enumerate(sample_ids):
        path = os.path.join(final_dir, "%06d.txt" % int(sid))
        if not os.path.exists(path):
            continue
        with open(path) as f:
            lines = [l.split() for l in f.read().splitlines() if l.strip()]
        if len(lines) > max_det:
            raise ValueError("%s holds %d detections, more than max_det=%d" % (path, len(lines), max_det))
        for k, parts in enumerate(lines):
            rec[i, k] = [float(v) for v in parts[3:3 + FIELDS]]
        cnt[i] = len(lines)
    return torch.from_numpy(rec), torch.from_numpy(cnt)


def write_result_dir(out_dir, sample_ids, records, counts, cls_name="Car"):
    """inverse of pack_result_dir: the line format of save_kitti_format (eval_rcnn.py:96-100)."""
    os.makedirs(out_dir, exist_ok=True)
    rec, cnt = records.cpu().numpy(), counts.cpu().numpy()
    for i, sid in enumerate(sample_ids):
        with open(os.path.join(out_dir, "%06d.txt" % int(sid)), "w") as f:
            for k in range(int(cnt[i])):
                print(("%s -1 -1" % cls_name) + "".join(" %.4f" % v for v in rec[i, k]), file=f)


def gather_results(records, counts, device=None):
    """THE collective: all ranks contribute (n_r, max_det, FIELDS) / (n_r,), padded to the largest
    shard, one all_gather each for the records and the counts -> per-rank lists on every rank."""
    world = dist.get_world_size()
    device = device or records.device
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    n_max = int(max(int(s.item()) for s in sizes))
    pad_rec = torch.zeros((n_max,) + tuple(records.shape[1:]), dtype=records.dtype, device=device)
    pad_cnt = torch.zeros((n_max,), dtype=counts.dtype, device=device)
    pad_rec[:records.shape[0]] = records.to(device)
    pad_cnt[:counts.shape[0]] = counts.to(device)
    all_rec = [torch.empty_like(pad_rec) for _ in range(world)]
    all_cnt = [torch.empty_like(pad_cnt) for _ in range(world)]
    dist.all_gather(all_rec, pad_rec)
    dist.all_gather(all_cnt, pad_cnt)
    return [(all_rec[r][:int(sizes[r].item())], all_cnt[r][:int(sizes[r].item())]) for r in range(world)]


def merge_sharded_results(all_ids, rank_final_dir, merged_dir, cls_name="Car", device=None):
    """pack this rank's files, all_gather, rank 0 writes `merged_dir`; returns the number of detections (rank 0)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard_ids(all_ids, rank, world)
    rec, cnt = pack_result_dir(rank_final_dir, mine)
    gathered = gather_results(rec, cnt, device=device)
    total = 0
    if rank == 0:
        for r, (rrec, rcnt) in enumerate(gathered):
            write_result_dir(merged_dir, shard_ids(all_ids, r, world), rrec, rcnt, cls_name)
            total += int(rcnt.sum().item())
    return total
