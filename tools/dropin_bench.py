#!/usr/bin/env python
"""The zero-change drop-in path, timed: the three calls unmodified main.py makes per frame
(/root/reference/main.py:57-59) on config 2 — yield_frame -> project_all_camera -> render_maps per camera on a blank
frame (the CPU baseline's convention; render_vectors adds the JPEG decode, which is host I/O) — one process, host
lists of NumPy arrays in and out, as the reference's protocol demands.

    python tools/dropin_bench.py [--frames 40] [--repeats 3]
"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def dropin_loop(cm, dataset, H, W, max_frames=None):
    """-> (seconds per phase dict, cam-frames, checksum)"""
    t = {"yield_frame": 0.0, "project_all_camera": 0.0, "render_maps": 0.0}
    done, check = 0, 0
    gen = cm.yield_frame(dataset)
    k = 0
    while max_frames is None or k < max_frames:
        t0 = time.perf_counter()
        try:
            image_idx, instance_map = next(gen)
        except StopIteration:
            break
        t1 = time.perf_counter()
        maps_2d = cm.project_all_camera(instance_map)
        t2 = time.perf_counter()
        for cam in cm.cm_list:
            image = cam.render_maps(np.zeros((H, W, 3), np.uint8), maps_2d[cam.camera_name])
            check += int(image[::7, ::7].sum())
            done += 1
        t3 = time.perf_counter()
        t["yield_frame"] += t1 - t0
        t["project_all_camera"] += t2 - t1
        t["render_maps"] += t3 - t2
        k += 1
    return t, done, check


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    import torch
    from cama_b200 import synth
    from cama_b200.dataset import ClipManager
    root = tempfile.mkdtemp()
    spec = synth.config2_spec(seed=0, name="dropin_config2")
    spec.write_cama = False
    clip = synth.write_clip(spec, root)
    cm = ClipManager(synth.CAMA_CONFIGS, clip, device=0, progress=False)
    dropin_loop(cm, "nuscenes", 540, 960, max_frames=3)            # warm-up: library load, scratch buffers
    best = None
    for _ in range(args.repeats):
        torch.cuda.synchronize()
        t, done, check = dropin_loop(cm, "nuscenes", 540, 960, max_frames=args.frames)
        total = sum(t.values())
        if best is None or total < best[0]:
            best = (total, t, done, check)
    total, t, done, check = best
    print(json.dumps({"workload": "config 2 through the per-frame drop-in calls (yield_frame, project_all_camera, render_maps on blank frames)",
                      "cam_frames": done, "seconds": round(total, 4), "cam_frames_per_s": round(done / total, 1),
                      "ms_per_frame": {k: round(1e3 * v / (done / 6), 3) for k, v in t.items()}, "checksum": check}))


if __name__ == "__main__":
    main()
