"""End-to-end evaluation of a KITTI-format tree at GPU speed: .bin files on disk -> KITTI result files, the
final_result/data output of `eval_rcnn.py --eval_mode rcnn` (pointrcnn/tools/eval_rcnn.py:466-649), with

  * the point pipeline of the dataset on the GPU (datasets/gpu_loader.py, csrc/scene_prepare.cu),
  * the batched, sync-free detector with several batches in flight (inference.Detector.submit / collect),
  * result files written by a small thread pool while the GPU works on the next batches.

    python tools/eval_fast.py --data_root <multi_data/kitti> --output_dir OUT [--batch_size 16] [--depth 3]
                              [--ckpt model.pth] [--per_scene_seed]

With --per_scene_seed (or a sharded run) the files are byte-identical to the unmodified eval_rcnn.py run with
PN2_PER_SCENE_SEED=1 (tests/test_gpu_loader_gpu.py); without it the np.random stream is consumed in scene order
exactly like eval_rcnn.py with --workers 0.  Under torchrun (one process per GPU) every rank evaluates
sample_id_list[rank::world] with per-scene seeds (the dataset shards itself, datasets/kitti_rcnn_dataset.py), writes
the result files of its own scenes into the shared output directory, and the ranks meet in ONE all_gather of their
(scenes, detections) counts before rank 0 adds the empty files: the directory is byte-identical to a single-GPU
--per_scene_seed run."""
import argparse
import concurrent.futures
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def load(sub):
    return importlib.import_module(PKG + "." + sub)


def run(data_root, output_dir, batch_size=16, depth=3, ckpt=None, per_scene_seed=False, seed=666, gpu_loader=True,
        split=None, writers=4, log=print):
    import torch
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        os.environ["PN2_SHARD_RANK"], os.environ["PN2_SHARD_WORLD"] = str(rank), str(world)   # the dataset keeps its shard
    cfgm = load("config")
    cfgm.use_default_yaml("rcnn")
    cfg = cfgm.cfg
    inf, ko = load("inference"), load("kitti_output")
    ds_mod, gl = load("datasets.kitti_rcnn_dataset"), load("datasets.gpu_loader")
    if per_scene_seed:
        os.environ["PN2_PER_SCENE_SEED"] = "1"
    dev = torch.device("cuda", torch.cuda.current_device())
    dataset = ds_mod.KittiRCNNDataset(root_dir=data_root, npoints=cfg.RPN.NUM_POINTS, split=split or cfg.TEST.SPLIT,
                                      mode='EVAL', random_select=True, classes=cfg.CLASSES)
    model = inf.build_model(seed=0, device=dev)
    if ckpt:
        load("train_utils").load_checkpoint(model, filename=ckpt)
    det = inf.Detector(model, dev, depth=depth)
    final_dir = os.path.join(output_dir, "final_result", "data")
    os.makedirs(final_dir, exist_ok=True)
    np.random.seed(seed)                                                   # eval_rcnn.py:467
    pool = concurrent.futures.ThreadPoolExecutor(max_workers=writers)
    jobs, n_det = [], [0]

    def write_batch(rec, cnt, sample_ids, shapes):
        for k, sid in enumerate(sample_ids):
            n = int(cnt[k])
            if n == 0:
                continue                                                   # eval_rcnn.py:616-617: no file for this scene
            calib = dataset.get_calib(int(sid))
            ko.save_kitti_format(int(sid), calib, rec[k, :n, :7], final_dir, rec[k, :n, 7], (shapes[k][0], shapes[k][1], 3))
            n_det[0] += n

    def batches():
        if gpu_loader:
            yield from gl.GpuSceneLoader(dataset, dev, batch_size=batch_size)
        else:                                                              # the reference's CPU data path, one process
            for start in range(0, len(dataset), batch_size):
                items = [dataset[i] for i in range(start, min(len(dataset), start + batch_size))]
                b = dataset.collate_batch(items)
                b["pts_input"] = torch.from_numpy(b["pts_input"]).pin_memory()
                b["img_shape"] = [dataset.get_image_shape(int(s))[:2] for s in b["sample_id"]]
                yield b

    t0 = time.perf_counter()
    pending = []

    def drain_one():
        ticket, ids, shapes = pending.pop(0)
        h_rec, h_cnt = det.collect(ticket)
        jobs.append(pool.submit(write_batch, h_rec.numpy().copy(), h_cnt.numpy().copy(), ids, shapes))

    n_scenes = 0
    for b in batches():
        if len(pending) == det.depth:
            drain_one()
        pending.append((det.submit(b["pts_input"], to_host=True), b["sample_id"], b["img_shape"]))
        n_scenes += len(b["sample_id"])
    while pending:
        drain_one()
    for j in jobs:
        j.result()
    dt = time.perf_counter() - t0
    split_file = os.path.abspath(os.path.join(dataset.imageset_dir, '..', '..', 'ImageSets', dataset.split + '.txt'))
    if dist is not None:
        # the single collective: every rank's (scenes, detections); it is also the point after which all result files exist
        mine = torch.tensor([n_scenes, n_det[0]], dtype=torch.int64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        n_scenes, n_det[0] = int(sum(int(t[0]) for t in every)), int(sum(int(t[1]) for t in every))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    empty = 0
    if rank == 0:
        empty = ko.dump_empty_files(final_dir, [x.strip() for x in open(split_file).readlines()])
    if rank == 0:
        log("eval_fast: %d scenes, %d detections, %d empty files, %.2f s = %.1f scenes/s (%s data path, %d batches in flight, "
            "%d GPU%s)" % (n_scenes, n_det[0], empty, dt, n_scenes / dt, "GPU" if gpu_loader else "CPU", depth, world,
                           "s" if world > 1 else ""))
    return {"scenes": n_scenes, "detections": n_det[0], "seconds": dt, "scenes_per_s": n_scenes / dt, "final_dir": final_dir}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data_root", required=True, help="<...>/multi_data/<dataset> (contains KITTI/)")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--batch_size", type=int, default=16)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--per_scene_seed", action="store_true")
    ap.add_argument("--cpu_loader", action="store_true", help="the reference's numpy data path (for comparison)")
    args = ap.parse_args()
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    res = run(args.data_root, args.output_dir, args.batch_size, args.depth, args.ckpt, args.per_scene_seed,
              gpu_loader=not args.cpu_loader)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({k: v for k, v in res.items()}))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
