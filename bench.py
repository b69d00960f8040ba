#!/usr/bin/env python
"""bench.py -- PointRCNN inference throughput (scenes/sec) on synthetic 16384-point KITTI-shaped
clouds; BASELINE.json's metric on its config 4 (full RPN+RCNN forward, batch 16, random-init
default.yaml weights) plus the eval post-processing of eval_rcnn.py:516-627.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one batch of 16 scenes through the whole hot path (H2D-free for `value`, from pinned
host buffers with the detections copied back for `e2e`).  Scenes shard across ranks (weak
scaling); the only collective is one all_gather of the detection records at the end of the e2e
region.  One JSON line on stdout (rank 0).

--impl reference runs the STOCK REFERENCE: its own, unmodified Python (lib/net/point_rcnn.py, pointnet2_lib/pointnet2/*.py,
lib/rpn/proposal_layer.py, the eval loop body of tools/eval_rcnn.py:497-627; oracle/refnet_gpu.py imports them from
baseline/_ref/pointrcnn) over its own CUDA kernels (oracle/_ref/libpn2_legacy.so = the reference .cu files compiled
unchanged), on every rank's GPU, same weights, same batches, same JSON keys.  `--impl mirror` is round 1's arm (the package's
module mirrors with fused=False on the legacy kernels), kept as a cross-check.  The CPU port (oracle/cpu_forward.py) is timed
next to it on a bounded sample: the reference has no CPU implementation of this path; see DESIGN.md "Reference arm".
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3d_adapt_auto_driving_b200"

METRIC = "pointrcnn_inference_scenes_per_sec"
UNIT = "scenes/s"
NPOINTS = 16384
FLOPS_PER_SCENE = 130.0e9  # 2*MAC of every SharedMLP layer of RPN + RCNN at default.yaml


def load(sub):
    return importlib.import_module(PKG + "." + sub)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
    if local < len(vis) and vis[local].strip().isdigit():
        return int(vis[local])
    return local


def make_batches(torch, syn, batch, nbatches, seed):
    return [torch.from_numpy(syn.make_clouds("lidar", batch, NPOINTS, seed=seed + 17 * i)).pin_memory()
            for i in range(nbatches)]


def dist_setup(torch, gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, world, rank, local
    torch.cuda.set_device(0)
    return None, 1, 0, 0


def cpu_port_sample(torch, batch_seed, max_scenes=4, budget_s=12.0):
    """The CPU port (oracle/cpu_forward.py) on a bounded sample of the same workload: scenes of
    the first benchmark batch, one at a time, all host threads for the torch part."""
    from oracle import cpu_forward as cf
    syn = load("synthetic")
    inf = load("inference")
    model = inf.build_model(seed=0, device="cpu")
    pkg = {"cfg": load("config").cfg, "decode_bbox_target": load("bbox_transform").decode_bbox_target}
    pts = torch.from_numpy(syn.make_clouds("lidar", max_scenes, NPOINTS, seed=batch_seed))
    t0 = time.perf_counter()
    done, dets = 0, []
    for i in range(max_scenes):
        out = cf.pointrcnn_forward(pkg, model, pts[i:i + 1])
        dets.append(cf.postprocess(pkg, out, 1)[0])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d scene(s) of batch 0 (16384 pts, 100 ROIs each), %.1f s; C restatement of the reference "
                      "kernels single-threaded + torch CPU fp32 convs on %d threads" % (done, dt, torch.get_num_threads()),
            "_detections": dets}


def match_detections(ref, got, tol=2e-3):
    """per scene lists of (boxes (k,7), scores (k,)) -> {"matched", "total", "extra", "max_abs_err"}: a reference box is
    matched when some produced box agrees with it in all seven fields and the raw score within `tol` (a score or an
    overlap within ~5e-5 of a threshold may legitimately flip a box: tests/test_refeval_golden_gpu.py)."""
    import numpy as np
    matched = total = extra = 0
    worst = 0.0
    for (rb, rs), (gb, gs) in zip(ref, got):
        total += len(rb)
        used = np.zeros(len(gb), bool)
        r = np.concatenate([np.asarray(rb, np.float64).reshape(-1, 7), np.asarray(rs, np.float64).reshape(-1, 1)], axis=1)
        g = np.concatenate([np.asarray(gb, np.float64).reshape(-1, 7), np.asarray(gs, np.float64).reshape(-1, 1)], axis=1)
        for row in r:
            if not len(g):
                break
            err = np.abs(g - row).max(axis=1)
            err[used] = np.inf
            j = int(err.argmin())
            if err[j] <= tol:
                used[j] = True
                matched += 1
                worst = max(worst, float(err[j]))
        extra += int((~used).sum())
    return {"matched": matched, "total": total, "extra": extra, "max_abs_err": worst, "tol": tol}


def workload_config(B, world):
    """the `config` both arms print (identical by construction)."""
    return {"workload": "BASELINE.json configs[3]: full PointRCNN RPN+RCNN forward (default.yaml, random-init weights "
                        "seed 0) + eval_rcnn.py decode/score/rotated-NMS, batch=16 synthetic KITTI-shaped clouds of "
                        "16384 points per GPU",
            "batch_per_gpu": B, "npoints": NPOINTS, "rois_per_scene": 100, "parallelism": "scene-shard x%d" % world,
            "l2": "per-step working set (pooled ROI tensor 0.44 GB + SA activations) >> 126 MB L2; inputs rotate over "
                  "4 batches"}


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    cabi = load("cabi")
    cabi.lib()  # fail loudly if the CUDA extension is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    dist, world, rank, local = dist_setup(torch, args.gpus)
    dev = torch.device("cuda", local)
    syn, inf = load("synthetic"), load("inference")
    model = inf.build_model(seed=0, device=dev)
    det = inf.Detector(model, dev, use_graph=not args.no_graph, depth=args.depth)
    B, K, W = args.batch, args.steps, args.warmup
    host = make_batches(torch, syn, B, 4, seed=1024 + 1000 * rank)
    resident = [h.to(dev) for h in host]
    out_rec = torch.empty((B, 100, 8), dtype=torch.float32).pin_memory()
    out_cnt = torch.empty((B,), dtype=torch.int32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_steps(n):
        """n steps with `depth` batches in flight (depth 1 = strictly one after the other)."""
        if args.depth <= 1:
            for i in range(n):
                det.detect_device(resident[i % 4])
        else:
            for i in range(n):
                det.submit(resident[i % 4])
            det.drain()

    run_steps(max(W, 2 * args.depth))
    torch.cuda.synchronize()
    if args.min_seconds > 0:
        # sustained figure: size K so that the timed region lasts at least --min-seconds (probe: 10 steps)
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); run_steps(10); e0.record(); torch.cuda.synchronize()
        K = max(K, int(args.min_seconds * 1e3 / (s0.elapsed_time(e0) / 10)) + 1)
        if dist is not None:
            t = torch.tensor([K], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            K = int(t.item())

    if args.minimal:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()     # ncu --profile-from-start off: only the timed steps are captured
        s.record()
        run_steps(K)
        e.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        emit({"minimal": True, "ms_per_step": s.elapsed_time(e) / K, "steps": K})
        return

    # which C-ABI kernel dominates a step (one profiled step, untimed)
    cabi.profile_start()
    with torch.no_grad():
        det._step(resident[0])
    breakdown = cabi.profile_stop()
    step_kernel_ms = sum(v["ms"] for v in breakdown.values())
    top = max(breakdown, key=lambda k: breakdown[k]["ms"])
    mlp_names = {"pn2_linear_f32", "pn2_sa_group_linear_f32", "pn2_linear_tc_f32", "pn2_linear_tc2_f32", "pn2_linear_pre_tc_f32",
                 "pn2_sa_group_linear_tc_f32", "pn2_sa_fused_tc_f32", "pn2_sa_fused_t_tc_f32", "pn2_rcnn_front_tc_f32"}
    # the shared-MLP kernels are one family (same contraction, three fusion levels): judged together
    mlp_ms = sum(v["ms"] for k, v in breakdown.items() if k in mlp_names)
    if mlp_ms >= breakdown[top]["ms"]:
        top = max((k for k in breakdown if k in mlp_names), key=lambda k: breakdown[k]["ms"])
    prof_names = mlp_names if top in mlp_names else {top}

    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()

    # ---- value: K steps, inputs resident in HBM ----
    barrier()
    l0 = cabi.launch_count
    if not det.use_graph:
        cabi.profile_start(prof_names)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run_steps(K)
    e.record()
    barrier()
    ms_total = max_over_ranks(s.elapsed_time(e))
    if det.use_graph:
        # the timed steps replay a CUDA graph, inside which single kernels cannot be bracketed by events:
        # the dominant-kernel duration is measured live on K more steps of the same work launched eagerly
        launches = det.launches_per_step
        cabi.profile_start(prof_names)
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        with torch.no_grad():
            for i in range(K):
                det._step(resident[i % 4])
        e2.record()
        torch.cuda.synchronize()
        dom = cabi.profile_stop()
        eager_ms = s2.elapsed_time(e2)
    else:
        dom = cabi.profile_stop()
        launches = (cabi.launch_count - l0) // K
        eager_ms = s.elapsed_time(e)

    # ---- e2e: host buffers in, detections out, every step; one all_gather at the end ----
    for i in range(2):
        det.detect(host[i % 4], out_rec, out_cnt)
    keep_rec = torch.empty((K, B, 100, 8), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, K, B, 100, 8), dtype=torch.float32, device=dev) if dist is not None else None
    if dist is not None:
        dist.all_gather_into_tensor(gathered, keep_rec)      # untimed: NCCL channel set-up for this size
    checksum = 0.0
    barrier()
    t0 = time.perf_counter()
    if args.depth <= 1:
        for i in range(K):
            pts = host[i % 4].to(dev, non_blocking=True)
            rec, num = det.detect_device(pts)
            keep_rec[i].copy_(rec)
            out_rec.copy_(rec, non_blocking=True)
            out_cnt.copy_(num, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the host reads the detections of every step
            checksum += float(out_cnt.sum())
    else:
        # the public pipelined API: H2D of a pinned batch, graph replay and D2H of its detections are enqueued
        # on the slot's stream; the host reads the detections of step i while step i+1 is on the GPU
        pending = []
        for i in range(K):
            if len(pending) == args.depth:
                h_rec, h_num = det.collect(pending.pop(0))
                checksum += float(h_num.sum())
            pending.append(det.submit(host[i % 4], to_host=True, keep=keep_rec[i]))
        while pending:
            h_rec, h_num = det.collect(pending.pop(0))
            checksum += float(h_num.sum())
        det.drain()
    if dist is not None:
        dist.all_gather_into_tensor(gathered, keep_rec)     # the single collective: detection boxes
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()

    scenes = B * K * world
    value = scenes / (ms_total * 1e-3)
    peaks = measured_peaks()
    dom_ms = sum(v["ms"] for v in dom.values())
    dom_work = sum(v["work"] for v in dom.values())
    dom_launches = sum(v["launches"] for v in dom.values())
    if top in mlp_names:
        # dense contraction with fp32 semantics: quoted against the TF32 dense tensor peak
        # (= half the measured bf16 figure; a 3xTF32 split needs 3 tensor FLOPs per useful one)
        peak = peaks["bf16_tflops_sustained"] / 2.0
        roof = {"bound": "tensor", "kernel": "+".join(sorted(dom)), "achieved": dom_work / (dom_ms * 1e-3) / 1e12,
                "peak": peak, "unit": "TFLOP/s", "peak_source": "%s bf16 sustained / 2 (TF32 dense)" % peaks["source"],
                "traffic": None}
    else:
        roof = {"bound": "hbm", "kernel": top, "achieved": dom_work / (dom_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "peak_source": peaks["source"], "traffic": None}
    # dram bytes of the same kernel family from the committed ncu --set full capture (per launch, like `achieved`)
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))     # newest capture last (r1, r1b, ...)
    tpath = tpaths[-1] if tpaths else ""
    if top in mlp_names and tpath:
        with open(tpath) as f:
            tj = json.load(f)
        roof["traffic"] = tj["mlp_family"]["dram_bytes_per_launch"]
        roof["traffic_source"] = tj["source"]
        roof["algorithmic_flop_per_launch"] = dom_work / max(dom_launches, 1)
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["launches_per_step"] = dom_launches // K
    roof["share_of_step"] = dom_ms / eager_ms      # of the same eagerly launched, one-at-a-time steps
    roof["timed_on"] = "K eager steps after the graph-replayed timed region" if det.use_graph else "the timed region"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world),
        "execution": {"launch": ("one CUDA graph replay per step" if det.use_graph else "eager")
                                + (", %d steps in flight on %d streams" % (args.depth, args.depth) if args.depth > 1 else ""),
                      "batches_in_flight": args.depth, "eager_ms_per_step": eager_ms / K,
                      "timed_region_s": ms_total * 1e-3,
                      "sampling": "bench batches are fixed synthetic clouds; dataset-level runs (tools/eval_sharded.py, "
                                  "tools/eval_fast.py) seed np.random per scene when sharded, a documented deviation from "
                                  "the reference's single global stream (DESIGN.md 6)"},
        "e2e": {"value": scenes / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * NPOINTS * 3 * 4,
                "d2h_bytes_per_step": B * 100 * 8 * 4 + B * 4,
                "collective": "one all_gather of (K,B,100,8) f32 detection records" if world > 1 else None,
                "detections_read_on_host": checksum},
        "gpu_launches": int(launches * K),
        "gpu_launches_per_step": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernel_breakdown_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])},
        "kernel_ms_per_step_sum": round(step_kernel_ms, 3),
        # FPS against the HBM roofline (north_star): scan bytes are SURVEY 8(d)'s 16 B per point and round of the reference's
        # scan, not DRAM traffic (the cloud lives on chip; since round 2 the pruned kernel does not even form most pairs:
        # the figure says how fast the reference's scan work is disposed of, like `culled_search` below).
        "scan_roofline": {k: {"ms": round(v["ms"], 4), "scan_GBps": round(v["work"] / (v["ms"] * 1e-3) / 1e9, 1),
                              "frac_of_hbm_peak": round(v["work"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 3)}
                          for k, v in breakdown.items() if k == "pn2_fps_cluster_f32" and v["ms"] > 0},
        # The default FPS kernel (csrc/fps_cells.cu) PRUNES: a round touches ~3 of 128 cells, so a fraction of the HBM peak over
        # the reference's scan bytes says as little as for the culled neighbour searches below (it comes out above 1).  Reported:
        # the rate at which the reference's scan bytes are disposed of, and the kernel's own unit, sampling rounds per second
        # and cloud (the launches of a step run B clouds side by side, one CTA each).
        "pruned_fps": {k: {"ms": round(v["ms"], 4), "reference_scan_GBps": round(v["work"] / (v["ms"] * 1e-3) / 1e9, 1),
                           "launches_per_step": v["launches"]}
                       for k, v in breakdown.items() if k in ("pn2_fps_f32", "pn2_fps_xyz_f32") and v["ms"] > 0},
        # The culled neighbour searches never form most centre-point pairs, so a bandwidth fraction over the reference's
        # scan bytes says nothing about them (it came out at 2.7 / 6.7 of the HBM peak).  Reported instead: the rate at
        # which the REFERENCE's pair tests are disposed of (B * M * N pairs of the brute-force kernels per second); the
        # limiter is occupancy / tail imbalance, not bandwidth (ncu: profiles/*_ncu_full_summary.csv, warps_active_pct).
        "culled_search": {k: {"ms": round(v["ms"], 4), "reference_pairs_per_s": round(v["work"] / 12.0 / (v["ms"] * 1e-3), 0)}
                          for k, v in breakdown.items()
                          if k in ("pn2_ball_query_culled_f32", "pn2_ball_query_culled_fill_f32", "pn2_ball_query_f32", "pn2_ball_query_dual_f32",
                                   "pn2_three_nn_culled_f32", "pn2_three_nn_f32") and v["ms"] > 0},
        "mlp_tflops_effective": FLOPS_PER_SCENE * scenes / (ms_total * 1e-3) / 1e12,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_sample(torch, 1024)
        ref_dets = cpu.pop("_detections")
        # parity inside the bench: the CPU port's detections for the first scenes of batch 0 against what the timed
        # configuration (B=16 x 16384, CUDA graph, batches in flight) produces for the same batch
        h_rec, h_cnt = det.detect(host[0], out_rec, out_cnt)
        got = inf.records_to_lists(h_rec, h_cnt)[:len(ref_dets)]
        line["parity_in_bench"] = dict(match_detections(ref_dets, got), scenes=len(ref_dets),
                                       against="oracle/cpu_forward.py (CPU port, pinned bit-for-bit to the reference's "
                                               "own network code: tests/test_refnet_vs_port_cpu.py)")
        line["cpu_baseline"] = cpu
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The stock reference (oracle/refnet_gpu.py: its unmodified Python over its unmodified kernels) on every rank."""
    import torch
    from oracle import refnet_gpu
    B, K, W = args.batch, args.steps, args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world)}
    if not refnet_gpu.available("legacy"):
        # no GPU / no legacy library / no staged reference tree: the CPU port on rank 0 is all there is
        if rank != 0:
            return
        cpu = cpu_port_sample(torch, 1024)
        cpu.pop("_detections")
        base.update({"n_gpus": 1, "value": cpu["value"], "ms_per_step": 1e3 * B / cpu["value"], "cpu_baseline": cpu,
                     "execution": {"arm": "CPU port (oracle/) of the same workload; stock reference unavailable here"},
                     "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
        return
    dist, world, rank, local = dist_setup(torch, args.gpus)
    dev = torch.device("cuda", local)
    syn, inf = load("synthetic"), load("inference")
    state = inf.build_model(seed=0, device="cpu").state_dict()         # the b200 arm's weights
    ref = refnet_gpu.Reference(state, dev, backend="legacy")
    host = make_batches(torch, syn, B, 4, seed=1024 + 1000 * rank)      # the b200 arm's batches
    sampler = ClockSampler(physical_gpu_index(local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        dets = ref.eval_batch(host[i % 4])
    sampler.start()
    barrier()
    copies = refnet_gpu.COPIED
    copies["h2d"] = copies["d2h"] = 0
    t0 = time.perf_counter()
    ndet = 0
    for i in range(K):
        dets = ref.eval_batch(host[i % 4])
        ndet += sum(len(d[1]) for d in dets)
    barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    clocks = sampler.stop()
    value = B * K * world / dt
    base.update({
        "value": value, "ms_per_step": 1e3 * dt / K,
        "execution": {"arm": "stock reference: unmodified lib/net/*.py, pointnet2_lib/pointnet2/*.py, lib/rpn/proposal_layer.py, "
                             "lib/utils/* and the eval-loop body of tools/eval_rcnn.py:497-627 (baseline/_ref/pointrcnn) over "
                             "oracle/_ref/libpn2_legacy.so (reference .cu compiled unchanged for sm_100a); cuDNN 1x1 convs at "
                             "torch defaults (cudnn.allow_tf32=%s); per-scene NMS with the blocking mask copy and host "
                             "greedy pass; one batch at a time on the default stream" % torch.backends.cudnn.allow_tf32,
                      "batches_in_flight": 1, "timed_region_s": dt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": copies["h2d"] // K, "d2h_bytes_per_step": copies["d2h"] // K,
                "detections_read_on_host": float(ndet)},
        "clocks": clocks,
    })
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_sample(torch, 1024)
        ref_dets = cpu.pop("_detections")
        got = ref.eval_batch(host[0])[:len(ref_dets)]
        base["parity_in_bench"] = dict(match_detections(ref_dets, got), scenes=len(ref_dets),
                                       against="oracle/cpu_forward.py (CPU port)")
        base["cpu_baseline"] = dict(cpu, note="the reference has no CPU implementation of this path; `value` above is its "
                                              "CUDA path on the same B200, this is the CPU port on the host cores")
    if rank == 0:
        emit(base)
    if dist is not None:
        dist.destroy_process_group()


def run_mirror(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, K, W = args.batch, args.steps, args.warmup
    from oracle import legacy
    have_gpu = torch.cuda.is_available() and legacy.available()
    cpu = cpu_port_sample(torch, 1024)
    cpu.pop("_detections")
    base = {"impl": "mirror", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if not have_gpu:
        base.update({"value": cpu["value"], "ms_per_step": 1e3 * B / cpu["value"], "cpu_baseline": cpu,
                     "config": {"workload": "CPU port (oracle/) of the same workload; legacy-CUDA library unavailable"},
                     "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    legacy.install(PKG)
    syn, inf = load("synthetic"), load("inference")
    cfg = load("config").cfg
    ku, iu = load("kitti_utils"), load("iou3d_utils")
    decode = load("bbox_transform").decode_bbox_target
    model = inf.build_model(seed=0, device=dev)
    for m in model.modules():
        if hasattr(m, "fused"):
            m.fused = False       # the reference's op-by-op composition (cuDNN convs, torch defaults)
    host = make_batches(torch, syn, B, 4, seed=1024)
    mean_size = torch.from_numpy(cfg.CLS_MEAN_SIZE[0]).to(dev)

    def step(pts_host):
        """eval_rcnn.py:498-629 for one batch."""
        with torch.no_grad():
            inputs = pts_host.cuda(non_blocking=True).float()
            ret = model({'pts_input': inputs})
            bs = inputs.shape[0]
            rois = ret['rois']
            rcnn_cls = ret['rcnn_cls'].view(bs, -1, ret['rcnn_cls'].shape[1])
            rcnn_reg = ret['rcnn_reg'].view(bs, -1, ret['rcnn_reg'].shape[1])
            pred = decode(rois.view(-1, 7), rcnn_reg.view(-1, rcnn_reg.shape[-1]), anchor_size=mean_size,
                          loc_scope=cfg.RCNN.LOC_SCOPE, loc_bin_size=cfg.RCNN.LOC_BIN_SIZE,
                          num_head_bin=cfg.RCNN.NUM_HEAD_BIN, get_xz_fine=True, get_y_by_bin=cfg.RCNN.LOC_Y_BY_BIN,
                          loc_y_scope=cfg.RCNN.LOC_Y_SCOPE, loc_y_bin_size=cfg.RCNN.LOC_Y_BIN_SIZE,
                          get_ry_fine=True).view(bs, -1, 7)
            raw = rcnn_cls
            norm = torch.sigmoid(raw)
            inds = norm > cfg.RCNN.SCORE_THRESH
            outs = []
            for k in range(bs):
                cur = inds[k].view(-1)
                if cur.sum() == 0:
                    continue
                bsel, ssel = pred[k, cur], raw[k, cur]
                keep = iu.nms_gpu(ku.boxes3d_to_bev_torch(bsel), ssel.view(-1), cfg.RCNN.NMS_THRESH).view(-1)
                outs.append((bsel[keep].cpu().numpy(), ssel[keep].cpu().numpy()))
            return outs

    for i in range(W):
        step(host[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        step(host[i % 4])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    value = B * K / dt
    base.update({
        "value": value, "ms_per_step": 1e3 * dt / K,
        "config": {"workload": "same as the b200 arm (configs[3], batch 16), executed by the REFERENCE implementation: "
                               "its CUDA kernels (oracle/_ref/libpn2_legacy.so, reference .cu compiled unchanged for sm_100a) "
                               "+ the reference's op-by-op module composition with cuDNN 1x1 convs (torch defaults, "
                               "cudnn.allow_tf32=%s) + per-scene host-greedy NMS, host buffers in / detections out"
                               % torch.backends.cudnn.allow_tf32, "batch_per_gpu": B, "npoints": NPOINTS},
        "cpu_baseline": dict(cpu, note="the reference has no CPU implementation of this path; `value` above is its "
                                       "CUDA path on the same B200, this is the CPU port on the host cores"),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })
    emit(base)


class QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." on the first
    collective), so file descriptor 1 points at stderr while the benchmark runs and is restored for the result line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def emit(line):
    """print the result line on the REAL stdout (see QuietStdout)."""
    sys.stdout.flush()
    os.write(QUIET.saved if QUIET is not None else 1, (json.dumps(line) + "\n").encode())


QUIET = None


def main():
    global QUIET
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "mirror"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=5,
                    help="batches in flight (Detector.submit/collect); 1 = one batch at a time on one stream")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--min-seconds", type=float, default=0.0,
                    help="grow --steps until the timed region lasts this long (sustained figure; default: exactly --steps)")
    ap.add_argument("--minimal", action="store_true",
                    help="warm-up + timed steps only (no profiled step, e2e or CPU legs): the command ncu wraps")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    with QuietStdout() as q:
        QUIET = q
        try:
            if args.impl == "reference":
                run_reference(args)
            elif args.impl == "mirror":
                run_mirror(args)
            else:
                run_b200(args)
        finally:
            QUIET = None


if __name__ == "__main__":
    main()
