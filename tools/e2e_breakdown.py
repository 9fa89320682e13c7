#!/usr/bin/env python
"""Where the milliseconds of Reproject.__call__ (sparse transfer) go, on the GPU box."""
import ctypes, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cama_b200 import synth, _native as N
from cama_b200.batched import Reproject

root = tempfile.mkdtemp()
spec = synth.config2_spec(); spec.write_cama = False
clip = synth.write_clip(spec, root)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
for _ in range(3): rp("nuscenes")
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
reps = 20
for _ in range(reps):
    t0 = time.perf_counter(); idx, w2c = rp.frame_poses("nuscenes"); tick("frame_poses (host pose seek + inverse)", t0)
    t0 = time.perf_counter(); w2c_dev = torch.from_numpy(w2c).to(rp.rt.device); tick("H2D poses", t0)
    res = rp.resident("nuscenes")
    t0 = time.perf_counter(); records, n, fmt = rp.renderer.render_overlay(res, w2c_dev); tick("render_overlay (GPU + stats read)", t0)
    cur, pal = rp._ov_host[0], res.palette
    t0 = time.perf_counter(); cur[:n].copy_(records[:n], non_blocking=True); tick("D2H records", t0)
    frames = rp._host_frames
    target = N.OverlayTarget(frames.ctypes.data, frames.shape[0], frames.shape[1], frames.shape[2], frames.shape[3], 0, None)
    t0 = time.perf_counter(); N.check(N.lib().cama_overlay_apply_host(cur.data_ptr(), n, fmt, None if pal is None else pal.ctypes.data, ctypes.byref(target), 3, 0)); tick("host erase", t0)
    t0 = time.perf_counter(); N.check(N.lib().cama_overlay_apply_host(cur.data_ptr(), n, fmt, None if pal is None else pal.ctypes.data, ctypes.byref(target), 2, 0)); tick("host draw", t0)
t0 = time.perf_counter()
for _ in range(reps): rp("nuscenes")
tick("whole call", t0)
for k, v in T.items(): print(f"{k:45s} {1e3 * v / reps:8.3f} ms")
print("records", n, "bytes each", cur.shape[1] * 4, "cores", os.cpu_count())
