#!/usr/bin/env python
"""Device-only timing of cama_clip_render on one workload: ms per step + per-phase ms (for A/B experiments on the GPU box).

    python tools/quick_bench.py [--workload config2] [--steps 100] [--tag NAME]
"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from cama_b200 import synth, _native as N
from cama_b200.batched import Reproject

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--tag", default="")
ap.add_argument("--sparse", action="store_true", help="sparse overlay output instead of dense frames")
ap.add_argument("--ramp", type=float, default=0.4, help="seconds of untimed clock ramp (0 under ncu)")
ap.add_argument("--graph", action="store_true", help="capture one step in a CUDA graph and replay it")
ap.add_argument("--frames", default="", help="lo:hi = render only this block of the clip's frames (a rank's share of a sharded site)")
args = ap.parse_args()
root = tempfile.mkdtemp()
clip, dataset = B.make_clip(args.workload, root, 0)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
rt, res = rp.rt, rp.resident(dataset)
idx, w2c = rp.frame_poses(dataset)
if args.frames:
    lo, hi = (int(v) for v in args.frames.split(":"))
    idx, w2c = idx[lo:hi], np.ascontiguousarray(w2c[lo:hi])
w2c_dev = torch.from_numpy(w2c).to(rt.device)
F, C = len(idx), rp.renderer.n_cams
frames = torch.empty((F, C, B.H, B.W, 3), dtype=torch.uint8, device=rt.device)
rp.renderer.render(res, w2c_dev, out=frames, check=True)
step = lambda: rp.renderer.render(res, w2c_dev, out=frames, check=False)
if args.sparse:                                   # sparse output only: lit-chunk records, no dense frames
    records, n_rec, fmt = rp.renderer.render_overlay(res, w2c_dev)
    count = torch.zeros(4, dtype=torch.int32, device=rt.device)
    ov = {"records_ptr": records.data_ptr(), "count_ptr": count.data_ptr(), "capacity": int(records.shape[0]), "fmt": fmt}
    step = lambda: rp.renderer.enqueue_overlay(res, w2c_dev, ov)
import time
host_us = None
if args.graph:
    for _ in range(3): step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            step()
    frames.zero_()
    eager_step, step = step, g.replay
else:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50 if args.ramp > 0 else 1): step()
    host_us = (time.perf_counter() - t0) / (50 if args.ramp > 0 else 1) * 1e6          # enqueue cost per step (the queue is not full yet)
    torch.cuda.synchronize()
t_end = time.perf_counter() + args.ramp
while time.perf_counter() < t_end:
    step(); torch.cuda.synchronize()
for _ in range(5): step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(args.steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
ph = [0, 0, 0, 0]
if not args.graph:
    rt.profile_enable(args.steps)
    for _ in range(args.steps): step()
    torch.cuda.synchronize()
    ph = rt.profile_read().mean(axis=0)
    rt.profile_enable(0)
print(json.dumps({"tag": args.tag, "graph": args.graph, "host_enqueue_us": host_us and round(host_us, 1), "workload": args.workload, "ms_per_step": round(ms, 5), "cam_frames_per_s": round(F * C / ms * 1e3),
                  "phase_us": {n: round(float(v) * 1e3, 2) for n, v in zip(N.PHASE_NAMES, ph)}, "checksum": int(count[0].item()) if args.sparse else int(frames[:, :, ::7, ::7].sum().item())}))
