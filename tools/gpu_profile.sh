#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one eager bench step with its device time,
# (2) full-set captures of the shared-MLP launches and of the scan kernels of one step, exported to CSV
# on the box (the .ncu-rep files are too large to bring back).
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
M='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor|sm__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__inst_executed.sum|sm__warps_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__warp_issue_stalled.*_per_warp_active.pct|Kernel Name|^"ID"'
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"linear_tc_kernel|sa_fused_tc_kernel|sa_fused_t_tc_kernel|rcnn_front_tc_kernel" --profile-from-start off -f -o /tmp/ncu/prof_mlp \
    python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/bench_ncu2.log 2>&1; echo "ncu mlp exit $?"
ncu -i /tmp/ncu/prof_mlp.ncu-rep --page raw --csv > /tmp/ncu/prof_mlp_raw.csv 2>/dev/null
python - <<'PY'
import csv, re
M = re.compile(r'gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__inst_executed.sum$|sm__warps_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum$|smsp__warp_issue_stalled_.*_per_warp_active.pct|Kernel Name|^ID$')
for name in ("prof_mlp", "prof_scan"):
    try:
        rows = list(csv.reader(open("/tmp/ncu/%s_raw.csv" % name)))
    except FileNotFoundError:
        continue
    keep = [i for i, h in enumerate(rows[0]) if M.search(h)]
    with open("gpurun_out/%s_summary.csv" % name, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in keep if i < len(r)])
    print(name, len(rows) - 2, "launches,", len(keep), "columns")
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fps_kernel|fps_cells_kernel|argsort_desc|unique_count_blocks|compact_blocks|fps_prefix|ball_query|three_nn|three_interpolate|nms_kernel|roipool3d|pairwise|gather_rows|spatial_order|unique_count_kernel|compact_kernel|decode_kernel|proposal_select|proposal_assemble|rcnn_post" --profile-from-start off -f -o /tmp/ncu/prof_scan \
    python bench.py --steps 1 --warmup 3 --minimal --no-graph --depth 1 > gpurun_out/bench_ncu3.log 2>&1; echo "ncu scan exit $?"
ncu -i /tmp/ncu/prof_scan.ncu-rep --page raw --csv > /tmp/ncu/prof_scan_raw.csv 2>/dev/null
python - <<'PY'
import csv, re
M = re.compile(r'gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__inst_executed.sum$|sm__warps_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum$|smsp__warp_issue_stalled_.*_per_warp_active.pct|Kernel Name|^ID$')
rows = list(csv.reader(open("/tmp/ncu/prof_scan_raw.csv")))
keep = [i for i, h in enumerate(rows[0]) if M.search(h)]
with open("gpurun_out/prof_scan_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep if i < len(r)])
print("prof_scan", len(rows) - 2, "launches")
PY
ls -la gpurun_out/
