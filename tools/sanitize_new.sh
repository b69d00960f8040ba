#!/bin/bash
# compute-sanitizer over the kernels added after tools/sanitize.sh was written: the pruned one-CTA FPS (fps_cells_kernel, every
# CTA size, small clouds: racecheck instruments every shared-memory access), the FPS kernels writing their centres, the
# ball-query fill variant (with hit counts), the interpolation with in-kernel weights, the two-launch group compaction (from the
# lists and from hit counts), the bitonic argsort, the two-warp FPS of many small clouds, the paired NMS launch and the three-phase
# ROI-pooling input stage.
# Logs -> gpurun_out/sanitizer_new_<tool>.log.   usage: tools/sanitize_new.sh [per-tool timeout s]
T=${1:-240}
mkdir -p gpurun_out
SUBSET="tests/test_pn2_ops_gpu.py tests/test_glue_gpu.py::test_argsort_desc_is_torch_sort tests/test_glue_gpu.py::test_nms_pair_launch_equals_two_launches tests/test_glue_gpu.py::test_rcnn_input_stage_one_launch_matches_torch_flow tests/test_linear_tc_gpu.py::test_sa_fused_t_skips_padded_duplicates_exactly"
KEXPR="(fps_cells and (130 or 64-64 or 2049 or 3000-3000)) or (fps_writes and (512-128 or 1-1-3)) or (fill_variant and (300 or 100)) or three_interpolate_from or grid_follows or (argsort and (777 or 5000 or 2-1 or 1-2)) or (skips_padded and 16-16-32) or (many_small and (ties-512 or uniform-128)) or group_compaction_from or nms_pair or rcnn_input_stage_one_launch"
for tool in memcheck synccheck racecheck; do
    extra=""
    K="$KEXPR"
    # racecheck cannot follow the mbarrier / tcgen05.commit chains of the tensor-core kernels (tools/sanitize.sh, racecheck_tc)
    # and its instrumentation outlasts their bounded waits: the compaction test runs under memcheck and synccheck only
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all" && K="($KEXPR) and not skips_padded"
    start=$(date +%s)
    PN2_SANITIZER=1 timeout $T compute-sanitizer --tool $tool $extra --print-limit 6 --log-file gpurun_out/sanitizer_new_$tool.raw \
        python -m pytest $SUBSET -q -m gpu -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_new_$tool.pytest 2>&1
    rc=$?
    {
        echo "# compute-sanitizer --tool $tool $extra ; python -m pytest $SUBSET -q -m gpu -k \"$K\""
        echo "# exit code $rc (124 = the $T s budget ran out before the subset finished), $(( $(date +%s) - start )) s"
        echo "# ---- pytest tail ----"
        tail -5 gpurun_out/sanitizer_new_$tool.pytest
        echo "# ---- sanitizer report (head) ----"
        grep -v "Host Frame" gpurun_out/sanitizer_new_$tool.raw | head -60 2>/dev/null
        echo "# ---- sanitizer report (tail) ----"
        tail -8 gpurun_out/sanitizer_new_$tool.raw 2>/dev/null
    } > gpurun_out/sanitizer_new_$tool.log
    rm -f gpurun_out/sanitizer_new_$tool.raw gpurun_out/sanitizer_new_$tool.pytest
    echo "sanitizer $tool rc=$rc"
done
