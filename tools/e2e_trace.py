#!/usr/bin/env python
"""Timeline of one Reproject.__call__ (sparse transfer) on the host clock: which step of the call takes what.
Wraps the native entry points and the torch calls the method makes; averages over repetitions."""
import ctypes, os, sys, tempfile, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cama_b200 import synth, _native as N
from cama_b200 import batched
from cama_b200.batched import Reproject

root = tempfile.mkdtemp()
spec = synth.config2_spec(); spec.write_cama = False
clip = synth.write_clip(spec, root)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
for _ in range(5): rp("nuscenes")
T = collections.OrderedDict()
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[label] = T.get(label, 0.0) + time.perf_counter() - t0
    setattr(obj, name, timed)
lib = N.lib()
class LibProxy:
    def __getattr__(self, k): return getattr(lib, k)
proxy = LibProxy()
for name, label in (("cama_clip_render", "cama_clip_render (enqueue)"), ("cama_clip_stats_read", "cama_clip_stats_read (sync + counters)"),
                    ("cama_overlay_apply_host", "cama_overlay_apply_host (blank + 3 draws)"), ("cama_clip_workspace_bytes", "workspace query")):
    fn = getattr(lib, name)
    def make(fn, label):
        def timed(*a):
            t0 = time.perf_counter(); r = fn(*a); T[label] = T.get(label, 0.0) + time.perf_counter() - t0; return r
        return timed
    setattr(proxy, name, make(fn, label))
N.lib = lambda: proxy
wrap(rp, "frame_poses", "frame_poses")
wrap(rp, "_poses_to_device", "poses -> device (pinned, async)")
ev_sync = torch.cuda.Event.synchronize
def timed_sync(self):
    t0 = time.perf_counter(); ev_sync(self); T["event waits (D2H slices)"] = T.get("event waits (D2H slices)", 0.0) + time.perf_counter() - t0
torch.cuda.Event.synchronize = timed_sync
reps = 30
t0 = time.perf_counter()
for _ in range(reps): rp("nuscenes")
total = time.perf_counter() - t0
acc = 0.0
for k, v in T.items():
    print(f"{k:45s} {1e3 * v / reps:7.3f} ms"); acc += v
print(f"{'(python glue, copies enqueue, rest)':45s} {1e3 * (total - acc) / reps:7.3f} ms")
print(f"{'whole call':45s} {1e3 * total / reps:7.3f} ms")
