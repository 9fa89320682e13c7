#!/usr/bin/env python
"""Experiment: how much do two clips overlap when rendered on two streams (upper bound for pipelining geometry under the raster)."""
import ctypes, os, sys, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from cama_b200 import synth, _native as N
from cama_b200.batched import Reproject

root = tempfile.mkdtemp()
clip, dataset = B.make_clip("config2", root, 0)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
rt, res = rp.rt, rp.resident(dataset)
idx, w2c = rp.frame_poses(dataset)
w2c_dev = torch.from_numpy(w2c).to(rt.device)
F, C = len(idx), rp.renderer.n_cams
r = rp.renderer
frames0 = torch.empty((F, C, B.H, B.W, 3), dtype=torch.uint8, device=rt.device)
r.render(res, w2c_dev, out=frames0, check=True)
cap = r.capacity.get((id(res), F), 0)
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = 100
streams = [torch.cuda.Stream(priority=0) for _ in range(n_streams)]
slots = []
for s in streams:
    fr = torch.empty_like(frames0)
    d = r._desc(res, w2c_dev, F, fr, None, "auto", cap, None)
    need = ctypes.c_size_t(); N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=rt.device)
    slots.append((s, fr, d, ws))
def run(k):
    s, fr, d, ws = slots[k % n_streams]
    N.check(N.lib().cama_clip_render(rt.ctx, ctypes.byref(d), rt.ptr(ws), ws.numel(), ctypes.c_void_p(s.cuda_stream)))
for k in range(20): run(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(K): run(k)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ok = all(torch.equal(fr, frames0) for _, fr, _, _ in slots)
print(json.dumps({"streams": n_streams, "ms_per_step": round(1e3 * dt / K, 5), "cam_frames_per_s": round(F * C * K / dt), "identical": ok}))
