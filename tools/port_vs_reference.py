#!/usr/bin/env python
"""Times the oracle port (oracle/cama_oracle.py) and the UNMODIFIED reference (/root/reference) side by
side, phase by phase, on the same config-2 clip — build container only (the reference does not travel).

    python tools/port_vs_reference.py [--frames 8] [--repeats 5] [--tolerance 0.05] [--images]

`bench.py --impl reference` and `cpu_baseline` time the port, because /root/reference does not exist on
the GPU box.  That stands in for the reference only if the port costs what the reference costs; this
script is the check: per phase (transform + crop = yield_frame, project_all_camera, render_maps on blank
frames, and with --images the imread + undistort-resize of render_vectors) it takes, per frame (per camera-frame for the image phases), the best of
`--repeats` interleaved runs of each implementation, sums them, prints the table and exits 1 when a phase of the
port is more than `--tolerance` slower or faster than the reference's.  Output kept under profiles/.
"""
import argparse
import os
import sys
import tempfile
import time
import types

sys.dont_write_bytecode = True
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--tolerance", type=float, default=0.05)
    ap.add_argument("--images", action="store_true", help="also time render_vectors' image path (writes 1600x900 JPEGs)")
    args = ap.parse_args()

    import cv2
    cv2.setNumThreads(1)
    import tqdm
    tqdm.tqdm = lambda it, *a, **k: it                         # the reference's progress bar (cama/dataset.py:88) off the clock
    import cama.dataset as ref_dataset                         # the reference
    ref_dataset.tqdm = lambda it, *a, **k: it
    from cama_b200 import synth
    from oracle import cama_oracle as orc

    with tempfile.TemporaryDirectory() as root:
        spec = synth.config2_spec(n_frames=args.frames, seed=0, name="pvr_config2")
        spec.write_cama = False
        clip = synth.write_clip(spec, root)
        if args.images:
            synth.write_background_jpegs(clip, args.frames, seed=0)
        cm = ref_dataset.ClipManager(synth.CAMA_CONFIGS, clip)
        oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
        h, w = orc.OUTPUT_HW

        def run_reference():
            t = {"yield_frame": [], "project_all_camera": [], "render_maps": [], "read_resized_image": []}
            check = 0
            gen = cm.yield_frame("nuscenes")
            while True:
                t0 = time.perf_counter()
                try:
                    image_idx, instance_map = next(gen)
                except StopIteration:
                    break
                t1 = time.perf_counter()
                maps_2d = cm.project_all_camera(instance_map)
                t2 = time.perf_counter()
                t["yield_frame"].append(t1 - t0)
                t["project_all_camera"].append(t2 - t1)
                for cam in cm.cm_list:
                    t3 = time.perf_counter()
                    image = cam.read_resized_image_by_index(image_idx) if args.images else np.zeros((h, w, 3), np.uint8)
                    t4 = time.perf_counter()
                    image = cam.render_maps(image, maps_2d[cam.camera_name])
                    t5 = time.perf_counter()
                    if args.images:
                        t["read_resized_image"].append(t4 - t3)
                    t["render_maps"].append(t5 - t4)
                    check += int(image[::7, ::7].sum())
            return t, check

        def run_port():
            t = {"yield_frame": [], "project_all_camera": [], "render_maps": [], "read_resized_image": []}
            check = 0
            gen = oc.frames("nuscenes")
            while True:
                t0 = time.perf_counter()
                try:
                    image_idx, chassis = next(gen)
                except StopIteration:
                    break
                t1 = time.perf_counter()
                per_cam = oc.project_all(chassis)
                t2 = time.perf_counter()
                t["yield_frame"].append(t1 - t0)
                t["project_all_camera"].append(t2 - t1)
                for cam in oc.cameras:
                    t3 = time.perf_counter()
                    image = oc.read_resized_image(cam, image_idx) if args.images else np.zeros((h, w, 3), np.uint8)
                    t4 = time.perf_counter()
                    image = orc.render_instances(image, per_cam[cam])
                    t5 = time.perf_counter()
                    if args.images:
                        t["read_resized_image"].append(t4 - t3)
                    t["render_maps"].append(t5 - t4)
                    check += int(image[::7, ::7].sum())
            return t, check

        best = {"reference": None, "port": None}
        checks = set()
        for r in range(args.repeats):
            for name, fn in (("reference", run_reference), ("port", run_port)) if r % 2 == 0 else (("port", run_port), ("reference", run_reference)):
                t, check = fn()
                checks.add(check)
                t = {k: np.array(v) for k, v in t.items()}
                # best of the repeats PER SAMPLE (one frame, or one camera-frame, of one phase): the build container's
                # cores are shared, and a whole-loop minimum still carries tens of per cent of interference
                best[name] = t if best[name] is None else {k: np.minimum(best[name][k], v) for k, v in t.items()}
        assert len(checks) == 1, f"port and reference rendered different images: {checks}"
        best = {name: {k: float(v.sum()) for k, v in t.items()} for name, t in best.items()}

    cam_frames = args.frames * len(oc.cameras)
    print(f"config 2, first {args.frames} frames x {len(oc.cameras)} cameras, one process, cv2 threads 1, per-sample best of {args.repeats} interleaved runs")
    print(f"{'phase':24s} {'reference s':>12s} {'port s':>12s} {'port/ref':>9s}")
    bad = []
    for phase in best["reference"]:
        a, b = best["reference"][phase], best["port"][phase]
        if a == 0.0 and b == 0.0:
            continue
        ratio = b / a
        print(f"{phase:24s} {a:12.4f} {b:12.4f} {ratio:9.3f}")
        if abs(ratio - 1.0) > args.tolerance:
            bad.append(phase)
    ta, tb = sum(best["reference"].values()), sum(best["port"].values())
    print(f"{'total':24s} {ta:12.4f} {tb:12.4f} {tb / ta:9.3f}")
    print(f"cam-frames/s: reference {cam_frames / ta:.1f}, port {cam_frames / tb:.1f}")
    if bad:
        print(f"FAIL: port differs from the reference by more than {args.tolerance:.0%} in: {', '.join(bad)}")
        return 1
    print(f"OK: every phase within {args.tolerance:.0%}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
