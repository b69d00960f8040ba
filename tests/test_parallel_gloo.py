"""The multi-GPU result exchange (parallel.py) on two CPU processes with the gloo backend:
per-rank KITTI result directories -> pack -> ONE all_gather -> rank 0 writes the merged directory,
byte-identical to the concatenation of the shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load, ROOT, PKG_NAME


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_results(final_dir, ids, seed):
    rng = np.random.RandomState(seed)
    os.makedirs(final_dir, exist_ok=True)
    for sid in ids:
        n = int(rng.randint(0, 6)) if sid % 5 else 0          # some scenes without detections
        with open(os.path.join(final_dir, "%06d.txt" % sid), "w") as f:
            for _ in range(n):
                vals = rng.uniform(-50, 50, 13)
                print("Car -1 -1" + "".join(" %.4f" % v for v in vals), file=f)


def _worker(rank, world, port, tmp, all_ids):
    import importlib, sys
    sys.path.insert(0, ROOT)
    par = importlib.import_module(PKG_NAME + ".parallel")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = par.shard_ids(all_ids, rank, world)
    final = os.path.join(tmp, "rank%d" % rank)
    _fake_results(final, mine, seed=100 + rank)
    # what the unmodified script does on every rank: empty files for scenes it does not own
    for sid in all_ids:
        p = os.path.join(final, "%06d.txt" % sid)
        if not os.path.exists(p):
            open(p, "w").close()
    total = par.merge_sharded_results(all_ids, final, os.path.join(tmp, "merged"))
    if rank == 0:
        with open(os.path.join(tmp, "total.txt"), "w") as f:
            f.write(str(total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_merge(tmp_path):
    world, all_ids = 2, list(range(0, 23))
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), all_ids), nprocs=world, join=True)
    par = load("parallel")
    n = 0
    for sid in all_ids:
        owner = sid % world
        want = open(os.path.join(str(tmp_path), "rank%d" % owner, "%06d.txt" % sid)).read()
        got = open(os.path.join(str(tmp_path), "merged", "%06d.txt" % sid)).read()
        assert got == want, sid
        n += len(want.splitlines())
    assert n > 0 and int(open(os.path.join(str(tmp_path), "total.txt")).read()) == n
    assert par.shard_ids(all_ids, 1, 2) == all_ids[1::2]


def test_pack_write_roundtrip(tmp_path):
    par = load("parallel")
    ids = [3, 4, 10]
    _fake_results(str(tmp_path / "a"), ids, seed=1)
    rec, cnt = par.pack_result_dir(str(tmp_path / "a"), ids)
    assert rec.shape == (3, par.MAX_DET, par.FIELDS) and rec.dtype == torch.float64
    par.write_result_dir(str(tmp_path / "b"), ids, rec, cnt)
    for sid in ids:
        assert open(str(tmp_path / "a" / ("%06d.txt" % sid))).read() == open(str(tmp_path / "b" / ("%06d.txt" % sid))).read()
