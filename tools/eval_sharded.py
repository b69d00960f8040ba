"""Run the reference's unmodified eval_rcnn.py scene-sharded over the GPUs of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/eval_sharded.py --tree <staged pointrcnn tree> --output_dir OUT -- <eval_rcnn.py arguments>

Every rank executes eval_rcnn.py (runpy, cwd = <tree>/tools, its own GPU, its own OUT/rank<r>) on
sample_id_list[rank::world]; then one all_gather of the detection records and rank 0 writes
OUT/merged/final_result/data (3d_adapt_auto_driving_b200/parallel.py)."""
import argparse
import glob
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", required=True, help="<dest>/pointrcnn made by evaltree.make_eval_tree")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    rest = [a for a in args.rest if a != "--"]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    os.environ["PN2_SHARD_RANK"], os.environ["PN2_SHARD_WORLD"] = str(rank), str(world)
    rank_out = os.path.join(os.path.abspath(args.output_dir), "rank%d" % rank)
    tools = os.path.join(args.tree, "tools")
    os.chdir(tools)
    sys.path.insert(0, tools)
    sys.argv = ["eval_rcnn.py"] + rest + ["--output_dir", rank_out]
    runpy.run_path(os.path.join(tools, "eval_rcnn.py"), run_name="__main__")

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    par = importlib.import_module(PKG + ".parallel")
    cfg = importlib.import_module(PKG + ".config").cfg
    dataset = rest[rest.index("--dataset") + 1] if "--dataset" in rest else "kitti"
    split_file = os.path.join(args.tree, "multi_data", dataset, "KITTI", "ImageSets", cfg.TEST.SPLIT + ".txt")
    all_ids = [int(x) for x in open(split_file).read().split()]
    finals = glob.glob(os.path.join(rank_out, "eval", "*", cfg.TEST.SPLIT, "**", "final_result", "data"), recursive=True)
    assert len(finals) == 1, finals
    merged = os.path.join(os.path.abspath(args.output_dir), "merged", "final_result", "data")
    total = par.merge_sharded_results(all_ids, finals[0], merged, cls_name=cfg.CLASSES, device=torch.device("cuda", local))
    if rank == 0:
        print("merged %d detections of %d scenes from %d ranks into %s" % (total, len(all_ids), world, merged))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
