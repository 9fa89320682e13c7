#!/usr/bin/env python
"""LiDAR aggregation (BASELINE.json configs[4]: 40 sweeps x 35 k points) on one GPU: device-resident timing of
cama_lidar_accumulate (+ zero-filling the count grid), the same through host arrays, and the NumPy oracle on the host."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cama_b200.lidar import LidarAggregator
from oracle import lidar_oracle
from test_lidar import _config5

sweeps, Ts = _config5()
agg = LidarAggregator()
rt = agg.rt
n_points = sum(len(s) for s in sweeps)
offsets = np.concatenate([[0], np.cumsum([len(s) for s in sweeps])]).astype(np.int64)
d_pts, d_off, d_T = rt.to_device(np.concatenate(sweeps)), rt.to_device(offsets), rt.to_device(Ts.reshape(-1, 16))
counts = agg.new_counts()
for _ in range(5):
    counts.zero_(); agg.accumulate_device(d_pts, d_off, d_T, counts, read_inside=False)
torch.cuda.synchronize()
K = 50
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
zero_ms = acc_ms = 0.0
for _ in range(K):
    e[0].record(); counts.zero_(); e[1].record(); agg.accumulate_device(d_pts, d_off, d_T, counts, read_inside=False); e[2].record()
    torch.cuda.synchronize()
    zero_ms += e[0].elapsed_time(e[1]); acc_ms += e[1].elapsed_time(e[2])
zero_ms /= K; acc_ms /= K
t0 = time.perf_counter()
for _ in range(5):
    c, inside = agg.accumulate(sweeps, Ts)
torch.cuda.synchronize()
host_ms = (time.perf_counter() - t0) / 5 * 1e3
t0 = time.perf_counter()
want, n_want = lidar_oracle.accumulate(sweeps, Ts, agg.origin, agg.voxel, agg.dims)
cpu_ms = (time.perf_counter() - t0) * 1e3
ok = bool(np.array_equal(c.cpu().numpy().view(np.uint32), want)) and inside == n_want
grid_bytes = counts.numel() * 4
print(json.dumps({"workload": "configs[4]: 40 sweeps x 35000 points (n,6) float64", "points": n_points, "inside": inside, "bit_exact_vs_oracle": ok,
                  "accumulate_ms": round(acc_ms, 4), "zero_grid_ms": round(zero_ms, 4), "grid_mb": grid_bytes / 1e6,
                  "points_per_s": round(n_points / (acc_ms * 1e-3)), "point_read_gbs": round(n_points * 48 / (acc_ms * 1e-3) / 1e9, 1),
                  "with_zeroing_points_per_s": round(n_points / ((acc_ms + zero_ms) * 1e-3)),
                  "host_arrays_ms": round(host_ms, 2), "numpy_oracle_ms": round(cpu_ms, 1)}))
