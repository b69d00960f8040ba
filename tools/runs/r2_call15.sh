#!/bin/bash
# round 2, call 15: cProfile of the unmodified script's main process; what-if (free pooling epilogue); RPN SA1 through the transposed compact kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python tools/profile_dropin.py > gpurun_out/r2c15_profile_dropin.log 2>&1; echo "profile rc=$?"
grep -n "cumulative\|tottime" gpurun_out/r2c15_profile_dropin.log | head -4
timeout 200 python tools/prof_sat.py > gpurun_out/r2c15_prof_sat.log 2>&1; echo "prof_sat rc=$?"; grep -E "ms \(incl|MMA warp total|EPI E|wait acc3 free" gpurun_out/r2c15_prof_sat.log | head -24
for v in 0 1 0 1; do
  PN2_SA_TRANSPOSED_SMALL=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c15_bench_small$v.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2c15_bench_small$v.json')); k=d['kernel_breakdown_ms_per_step']; print('transposed_small $v:', round(d['value'],1), round(d['ms_per_step'],3), 'sa_t', k.get('pn2_sa_fused_t_tc_f32'), 'sa', k.get('pn2_sa_fused_tc_f32'), 'compact', k.get('pn2_group_compact_i32'), d.get('parity_in_bench'))"
done
