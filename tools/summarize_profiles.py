#!/usr/bin/env python
"""gpurun_out/{launches.csv, prof_mlp_summary.csv, prof_scan_summary.csv} (written on the GPU box by
tools/gpu_profile.sh) -> the tracked evidence under profiles/:

    <tag>_ncu_launch_shares.csv   every kernel of one eager bench step: launches, total us, share
    <tag>_ncu_full_summary.csv    ncu --set full rows of our kernels (time, DRAM bytes, tensor pipe, ...)
    <tag>_traffic.json            DRAM bytes per launch of every kernel family (read by bench.py -> roofline.traffic)

    python tools/summarize_profiles.py r1b
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
OWN = re.compile(r"fps_kernel|fps_cells_kernel|argsort_desc|unique_count_blocks|compact_blocks|ball_query|three_nn|three_interpolate|nms_kernel|roipool3d|pairwise_kernel|gather_rows|"
                 r"scatter_rows|spatial_order|unique_count_kernel|compact_kernel|scene_filter|scene_gather|stat_rescale|d3_overlap|linear_tc_kernel|sa_fused_tc_kernel|sa_fused_t_tc_kernel|linear_kernel|rotate_iou|"
                 r"rcnn_front_tc_kernel|decode_kernel|proposal_select_kernel|proposal_assemble_kernel|rcnn_post_prepare_kernel|rcnn_post_assemble_kernel|roipool3d_canon_kernel|fps_prefix")
MLP = re.compile(r"linear_tc_kernel|sa_fused_tc_kernel|sa_fused_t_tc_kernel|linear_kernel|rcnn_front_tc_kernel")


def short(name):
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name)       # drop the argument list
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("at::", "")
    return name.replace(",", ";")[:70]


def unit_scale(unit, want):
    table = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
             "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return table[unit] / (1e6 if want == "MB" else 1.0)


def launch_shares(tag):
    path = os.path.join(OUT, "launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, body = rows[0], rows[1:]
    kn, mv = head.index("Kernel Name"), head.index("Metric Value")
    agg = OrderedDict()
    for r in body:
        d = agg.setdefault(short(r[kn]), [0, 0.0])
        d[0] += 1
        d[1] += float(r[mv].replace(",", "")) / 1e3
    total = sum(v[1] for v in agg.values())
    own = [(k, v) for k, v in agg.items() if OWN.search(k)]
    with open(os.path.join(PROF, tag + "_ncu_launch_shares.csv"), "w") as f:
        f.write("# %s ncu launch list, one eager bench step (B=16): `ncu --profile-from-start off --metrics "
                "gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --minimal --no-graph "
                "--depth 1`\n" % tag)
        f.write("# cold-cache, serialised per-launch times: compare SHARES with bench.py's kernel_breakdown, not absolutes\n")
        f.write("# launches in the step: %d, total %.1f us; own kernels: %d launches, %.1f us (%.1f%%)\n"
                % (len(body), total, sum(v[0] for _, v in own), sum(v[1] for _, v in own),
                   100.0 * sum(v[1] for _, v in own) / total))
        f.write("kernel,launches,total_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / total < 0.002 and not OWN.search(k):
                continue
            f.write("%s,%d,%.1f,%.2f\n" % (k, v[0], v[1], 100.0 * v[1] / total))
    return agg


def full_summary(tag):
    out_rows, fam = [], OrderedDict()
    for cap in ("prof_mlp", "prof_scan"):
        path = os.path.join(OUT, cap + "_summary.csv")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        head, units, body = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(head)}

        def get(r, name, want=None):
            if name not in col or col[name] >= len(r) or r[col[name]] == "":
                return None
            v = float(r[col[name]].replace(",", ""))
            return v * unit_scale(units[col[name]], want) if want else v
        for r in body:
            name = short(r[col["Kernel Name"]])
            ms = get(r, "gpu__time_duration.sum", "ms")
            rd, wr = get(r, "dram__bytes_read.sum", "MB"), get(r, "dram__bytes_write.sum", "MB")
            tp = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            out_rows.append([cap, name, "%.4f" % ms, "%.2f" % rd, "%.2f" % wr, "%.0f" % ((rd + wr) / ms),
                             "" if tp is None else "%.2f" % tp, "%d" % get(r, "launch__grid_size"),
                             "%d" % get(r, "launch__block_size"), "%d" % get(r, "launch__registers_per_thread"),
                             "%.0f" % get(r, "smsp__inst_executed.sum"),
                             "%.2f" % get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                             "%.2f" % get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                             "%.2f" % get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")])
            d = fam.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
            d["launches"] += 1
            d["ms"] += ms
            d["dram_read_MB"] += rd
            d["dram_write_MB"] += wr
    with open(os.path.join(PROF, tag + "_ncu_full_summary.csv"), "w") as f:
        f.write("# %s ncu --set full captures of one eager bench step (B=16, no graph, one batch at a time), exported on the box with\n"
                "#   ncu -i <rep> --page raw --csv ; columns reduced (tools/gpu_profile.sh, tools/summarize_profiles.py).\n"
                "#   Times are under the profiler (cold cache, serialised): shares and per-kernel diagnosis, never bench values.\n" % tag)
        f.write("capture,kernel,duration_ms,dram_read_MB,dram_write_MB,dram_GBps,tensor_pipe_active_pct,grid,block,regs,"
                "warp_inst,dram_pct_of_peak,sm_throughput_pct,warps_active_pct\n")
        for r in out_rows:
            f.write(",".join(r) + "\n")
    mlp = [v for k, v in fam.items() if MLP.search(k)]
    n_mlp = sum(v["launches"] for v in mlp)
    mlp_bytes = sum(v["dram_read_MB"] + v["dram_write_MB"] for v in mlp) * 1e6
    for v in fam.values():
        for k in ("ms", "dram_read_MB", "dram_write_MB"):
            v[k] = round(v[k], 4 if k == "ms" else 1)
    traffic = {"source": "profiles/%s_ncu_full_summary.csv (ncu --set full, one eager step, B=16)" % tag,
               "mlp_family": {"launches": n_mlp, "dram_bytes_per_step": mlp_bytes,
                              "dram_bytes_per_launch": mlp_bytes / max(n_mlp, 1)},
               "kernels": fam}
    with open(os.path.join(PROF, tag + "_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    return fam


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    agg = launch_shares(tag)
    fam = full_summary(tag)
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        print("%-40s n=%-3d %8.3f ms  rd %9.1f MB  wr %9.1f MB" % (k, v["launches"], v["ms"], v["dram_read_MB"], v["dram_write_MB"]))
