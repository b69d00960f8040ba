#!/bin/bash
# FPS v2 (rank-sorted registers, packed fp32, CTA-scope wait), batched meta warp, 4 CTAs/SM persistent grids,
# two batches in flight: parity first, then micro timings, the fused-SA stopwatch and the bench at depth 1 / 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu23.log
timeout 300 python - <<'PY' 2>&1 | tail -12
import importlib, sys, json
sys.path.insert(0, ".")
sys.argv = ["microbench"]
import torch, numpy as np
PKG = "3d_adapt_auto_driving_b200"
cabi = importlib.import_module(PKG + ".cabi"); syn = importlib.import_module(PKG + ".synthetic")
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))
for b, n, m in [(8, 16384, 4096), (16, 16384, 4096), (16, 4096, 1024), (16, 1024, 256), (1600, 512, 128), (1600, 128, 32)]:
    xyz = torch.from_numpy(syn.make_clouds("lidar", min(b, 16), n, seed=1024)).to(dev)
    if b > 16: xyz = xyz.repeat(b // 16, 1, 1).contiguous()
    idx = torch.empty((b, m), dtype=torch.int32, device=dev)
    out = {}
    for c in ([0] if n < 2048 else [0, 2, 4, 8]):
        if c and b * c > 148: continue
        cabi.lib().pn2_fps_set_cluster(c)
        ms = timeit(lambda: cabi.call("pn2_fps_f32", cabi.ptr(xyz), cabi.ptr(None), cabi.ptr(idx), cabi.i32(b), cabi.i32(n), cabi.i32(m)))
        out[c] = round(ms, 4)
    cabi.lib().pn2_fps_set_cluster(0)
    print("fps b=%d n=%d m=%d ms by cluster %s  scan GB/s (auto) %.0f" % (b, n, m, out, b * (m - 1) * n * 16 / out[0] / 1e6))
PY
timeout 300 python tools/prof_tc.py 2 2>&1 | tail -26
for d in 1 2 3; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/bench23_d$d.json 2> gpurun_out/bench23_d$d.err; echo "bench depth $d exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench23_d$d.json"))
print(d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["clocks"])
print(d["kernel_breakdown_ms_per_step"], d["kernel_ms_per_step_sum"])
PY
tail -3 gpurun_out/bench23_d$d.err
done
