#!/usr/bin/env python
"""Per-CTA timeline of the raster kernel (CAMA_RASTER_DEBUG=32): when do CTAs finish, how many active bands each processed."""
import ctypes, os, sys, tempfile
os.environ["CAMA_RASTER_DEBUG"] = str(int(os.environ.get("CAMA_RASTER_DEBUG", "0")) | 32)
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
frames = torch.empty((len(idx), 6, B.H, B.W, 3), dtype=torch.uint8, device=rt.device)
for _ in range(30): rp.renderer.render(res, w2c_dev, out=frames, check=False)
torch.cuda.synchronize()
n = rt.sm_count() * int(os.environ.get("CAMA_RASTER_CTAS", "4"))
buf = np.zeros((n, 3), dtype=np.uint64)
fn = N.lib().cama_debug_raster_timeline
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, n) == 0
t0 = buf[:, 0].min()
start, end, bands = (buf[:, 0] - t0) / 1e3, (buf[:, 1] - t0) / 1e3, buf[:, 2]
print(f"CTAs {n}: start us min/mean/max {start.min():.1f}/{start.mean():.1f}/{start.max():.1f}")
print(f"end us: min {end.min():.1f} p10 {np.percentile(end,10):.1f} median {np.median(end):.1f} mean {end.mean():.1f} p90 {np.percentile(end,90):.1f} max {end.max():.1f}")
print(f"active bands per CTA: min {bands.min()} mean {bands.mean():.2f} max {bands.max()}  total {bands.sum()}")
dur = end - start
print(f"us per active band (CTA mean): {np.mean(dur / np.maximum(bands, 1)):.2f}")
h, edges = np.histogram(end, bins=12)
print("end-time histogram:", [f"{e:.0f}:{c}" for e, c in zip(edges[:-1], h)])
