#!/bin/bash
# round 2, call 23: ncu source-level stall samples of the pruned FPS kernel v3 (8 and 16 warps)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out /tmp/ncu
cat > /tmp/one.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
cabi = importlib.import_module('3d_adapt_auto_driving_b200.cabi'); syn = importlib.import_module('3d_adapt_auto_driving_b200.synthetic')
xyz = torch.from_numpy(syn.make_clouds('lidar', 16, 16384, seed=1024)).cuda()
idx = torch.empty((16, 4096), dtype=torch.int32, device='cuda')
for w in (8, 8, 16):
    cabi.call('pn2_fps_cells_f32', cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(16), cabi.i32(16384), cabi.i32(4096), cabi.i32(w))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fps_cells -s 1 -c 2 -f -o /tmp/ncu/fps_cells python /tmp/one.py > gpurun_out/r2d_ncu_fps_cells_v3.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu/fps_cells.ncu-rep --page source --csv > gpurun_out/r2d_ncu_fps_cells_v3_source.csv 2>/dev/null
ncu -i /tmp/ncu/fps_cells.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_fps_cells_v3_raw.csv 2>/dev/null
ls -la gpurun_out/r2d_ncu_fps_cells_v3*
