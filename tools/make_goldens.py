"""Produce golden vectors from the REFERENCE implementations on the GPU box.

Run under gpurun (needs a GPU):  python tools/make_goldens.py gpurun_out/golden
  - legacy CUDA kernels (oracle/_ref/libpn2_legacy.so = the reference .cu files compiled
    unchanged): iou3d overlap / iou / NMS masks, roipool3d, FPS / ball_query / three_nn
  - evaluate/rotate_iou.py (numba.cuda), imported from the git-ignored copy that
    oracle/build_ref.py places at oracle/_ref/rotate_iou.py; also dumps the PTX numba/NVVM
    generated, which is the only place the kernel's real arithmetic (f64 sub-expressions,
    FMA contraction) is visible.
The resulting small .npz files are committed under tests/golden/ with this script.
"""
import hashlib
import json
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rotate_iou_inputs(seed, n):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-5, 5, size=(n, 2))
    d = rng.uniform(1, 4, size=(n, 2))
    a = rng.uniform(-np.pi, np.pi, size=(n, 1))
    return np.concatenate([c, d, a], 1).astype(np.float32)


def rotate_iou_adversarial():
    b = [
        [0, 0, 2, 2, 0], [0, 0, 2, 2, 0],              # identical
        [2, 0, 2, 2, 0],                                # shares an edge with the first
        [0, 0, 2, 2, np.pi / 2], [0, 0, 2, 2, -np.pi / 2], [0, 0, 4, 1, np.pi / 4],
        [0, 0, 0, 0, 0], [1, 1, 0, 3, 0.3],             # zero area
        [0.5, 0.5, 1, 1, 0], [0, 0, 1, 1, 0],           # contained / corner touching
        [10, 10, 1, 1, 1.0],                            # disjoint
        [0, 0, 2, 2, 1e-7], [0, 0, 2, 2, np.pi], [1e-3, 0, 2, 2, 0],
        [0, 0, 3.9, 1.6, 0.3], [0.2, 0.1, 3.9, 1.6, 0.31],
    ]
    return np.asarray(b, np.float32)


def bev_boxes(seed, n, spread=20.0):
    rng = np.random.RandomState(seed)
    cx = rng.uniform(-spread, spread, n); cz = rng.uniform(0, 2 * spread, n)
    l = rng.uniform(3.0, 4.8, n); w = rng.uniform(1.4, 2.0, n)
    ry = rng.uniform(-np.pi, np.pi, n)
    # many near-duplicates so NMS has work to do
    dup = rng.randint(0, n, n // 2)
    cx[: n // 2] = cx[dup] + rng.normal(0, 0.15, n // 2)
    cz[: n // 2] = cz[dup] + rng.normal(0, 0.15, n // 2)
    ry[: n // 2] = ry[dup] + rng.normal(0, 0.05, n // 2)
    return np.stack([cx - l / 2, cz - w / 2, cx + l / 2, cz + w / 2, ry], 1).astype(np.float32)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    report = {}
    import torch
    from oracle import legacy
    syn = __import__("importlib").import_module("3d_adapt_auto_driving_b200.synthetic")
    dev = torch.device("cuda:0")

    # ---------------- legacy iou3d ----------------
    a = bev_boxes(0, 300); b = bev_boxes(1, 200)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    ov = legacy.boxes_overlap_bev(ta, tb).cpu().numpy()
    iou = legacy.boxes_iou_bev(ta, tb).cpu().numpy()
    nb = bev_boxes(2, 1000)
    tnb = torch.from_numpy(nb).to(dev)
    keep = {}
    for thr in (0.1, 0.8):
        keep["rot_%g" % thr] = legacy.greedy_from_mask(legacy.nms_mask(tnb, thr, normal=False).cpu(), 1000)
        keep["nrm_%g" % thr] = legacy.greedy_from_mask(legacy.nms_mask(tnb, thr, normal=True).cpu(), 1000)
    np.savez_compressed(os.path.join(out_dir, "iou3d_legacy.npz"), a=a, b=b, overlap=ov, iou=iou, nms_boxes=nb,
                        **{"keep_" + k: v for k, v in keep.items()})
    report["iou3d"] = {k: int(len(v)) for k, v in keep.items()}

    # ---------------- legacy roipool3d ----------------
    xyz = syn.make_clouds("lidar", 2, 16384, seed=1024)
    rng = np.random.RandomState(3)
    feat = rng.randn(2, 16384, 5).astype(np.float32)
    boxes = np.zeros((2, 24, 7), np.float32)
    for bi in range(2):
        for m in range(20):
            p = xyz[bi, rng.randint(0, 16384)]
            boxes[bi, m] = [p[0], p[1] + 0.8, p[2], 1.5 + 2.0, 1.6 + 2.0, 3.9 + 2.0, rng.uniform(-np.pi, np.pi)]
        # the remaining 4 rows stay all-zero boxes (zero-padded ROIs still flow through pooling)
    pooled, empty = legacy.roipool3d(torch.from_numpy(xyz).to(dev), torch.from_numpy(feat).to(dev),
                                     torch.from_numpy(boxes).to(dev), sampled=512)
    pooled = pooled.cpu().numpy(); empty = empty.cpu().numpy()
    # store the pooled xyz+features only through their source indices (small) + a checksum
    np.savez_compressed(os.path.join(out_dir, "roipool3d_legacy.npz"), seed=1024, boxes=boxes, feat_seed=3,
                        empty=empty, pooled_xyz=pooled[..., :3], sha=np.frombuffer(hashlib.sha256(pooled.tobytes()).digest(), np.uint8))
    report["roipool3d"] = {"empty": int(empty.sum()), "sha": hashlib.sha256(pooled.tobytes()).hexdigest()}

    # ---------------- legacy pointnet2 (small fixtures for the CPU suite) ----------------
    g = {}
    for kind in ("uniform", "lidar", "ties"):
        x = syn.make_clouds(kind, 2, 4096, seed=1024)
        tx = torch.from_numpy(x).to(dev)
        idx, temp = legacy.fps(tx, 1024)
        new_xyz = torch.gather(tx, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        g[kind + "_fps_idx"] = idx.cpu().numpy()
        g[kind + "_bq_0.5_16"] = legacy.ball_query(0.5, 16, tx, new_xyz).cpu().numpy().astype(np.int16)
        d2, i3 = legacy.three_nn(tx, new_xyz)
        g[kind + "_nn_idx"] = i3.cpu().numpy().astype(np.int16)
        g[kind + "_nn_d2_sha"] = np.frombuffer(hashlib.sha256(d2.cpu().numpy().tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(out_dir, "pointnet2_legacy.npz"), **g)

    # ---------------- numba rotate_iou ----------------
    try:
        os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import numba
        from numba import cuda
        report["numba"] = {"version": numba.__version__}
        import rotate_iou as ref  # the unmodified reference file
        out = {}
        big_a, big_b = rotate_iou_inputs(0, 1000), rotate_iou_inputs(1, 1000)
        adv = rotate_iou_adversarial()
        small_a, small_b = big_a[:160], big_b[:130]
        for crit in (-1, 0, 1, 2):
            out["small_c%d" % crit] = ref.rotate_iou_gpu_eval(small_a, small_b, crit)
            out["adv_c%d" % crit] = ref.rotate_iou_gpu_eval(adv, adv, crit)
            big = ref.rotate_iou_gpu_eval(big_a, big_b, crit)
            out["big_sha_c%d" % crit] = np.frombuffer(hashlib.sha256(big.tobytes()).digest(), np.uint8)
            out["big_diag_c%d" % crit] = big[::7, ::11].copy()
        np.savez_compressed(os.path.join(out_dir, "rotate_iou_numba.npz"), adv=adv, **out)
        asm = ref.rotate_iou_kernel_eval.inspect_asm()
        for i, (sig, ptx) in enumerate(asm.items()):
            with open(os.path.join(out_dir, "rotate_iou_numba_%d.ptx" % i), "w") as f:
                f.write("// signature: %s\n" % (sig,))
                f.write(ptx)
        report["numba"]["ok"] = True
    except Exception:
        report.setdefault("numba", {})["error"] = traceback.format_exc()
    with open(os.path.join(out_dir, "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1)[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
