#!/usr/bin/env python
"""Benchmark of the reprojection hot path (BASELINE.json metric: reprojected camera-frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload config2]

A step = one pass of the hot path over one clip: every frame x camera x dense map vertex
transformed, cropped, projected, masked and stamped into the uint8 [F,C,540,960,3] frames
(reference loop: cama/dataset.py:78-126).  Workload at N=1 = BASELINE.json configs[1]
(40 frames x 6 cameras, 200 polylines, 0.1 m densify => ~96 k vertices).  At N>1 the workload is
BASELINE.json configs[3]: the config-3 site (320 frames x 6 cameras, ~767 k vertices) SHARDED BY FRAME over
the N GPUs (strong scaling), and a step ends with every frame of the site assembled in the HBM of every
rank — the assembly is inside the timed region.  The ranks exchange records through peer memory
(cama_b200/shard.py, csrc/peer.cu: up to 4 GPUs the centre records, after which every rank rasters every
frame; beyond, the lit 8-pixel chunks, mirrored by the raster while it runs and expanded into zero-filled
frames) instead of all-gathering the dense frames; the literal NCCL all-gather of the uint8 frames is timed
beside it (`dense_allgather`), as are the compute-only figure and one GPU rendering the same site alone,
and every rank checks its assembled frames byte for byte against its own single-GPU render (`verified`).

One JSON line on rank 0:
  value        cam-frames/s, inputs (vertices, poses) resident in HBM, frames left in HBM; the K steps are dealt
               over --lanes CUDA streams (default 3: independent clips overlap); single_stream = one stream
  e2e          same metric through Reproject.__call__: host pose lookup + float32 inverse, H2D of
               the poses, render, lit-chunk records back over PCIe and drawn into the host frames
               (e2e.dense: every frame byte copied back instead)
  roofline     the raster kernel (writes every frame byte once) against the measured HBM peak
  config3      the site of BASELINE.json configs[2] on this GPU: ms per site, raster and whole-step roofline
  dropin       the per-frame calls unmodified main.py makes (yield_frame, project_all_camera, render_maps)
  cpu_baseline the NumPy/OpenCV oracle (= the reference's loop) on this box's host cores (+ with_images)
`--impl reference` times that CPU path alone and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "reprojected camera-frames/sec"
UNIT = "cam-frames/s"
H, W = 540, 960


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["config2", "config2_cama", "config3"],
                    help="default: config2 on one GPU, config3 (the site, sharded by frame) on several")
    ap.add_argument("--mode", default="auto", choices=["auto", "binned", "plane"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the config-3 (site, 320 frames) block of the single-GPU line")
    ap.add_argument("--lanes", type=int, default=3, help="CUDA streams the K steps are dealt over (independent clips overlap); 1 = one stream")
    ap.add_argument("--geometry-ctas", type=int, default=3,
                    help="resident geometry CTAs per SM while clips overlap on several streams (cama_clip_desc.geometry_ctas_per_sm; 4 takes every "
                         "register of an SM, 3 lets raster CTAs of another clip in); the single-stream loops use the library default")
    ap.add_argument("--raster-ctas", type=int, default=3,
                    help="cama_clip_desc.raster_ctas_per_sm for the overlapping clips (0 = library default of 4: fastest for one clip alone; with 3 + 3 "
                         "CTAs per SM geometry and raster CTAs of different clips co-reside)")
    ap.add_argument("--ramp-seconds", type=float, default=0.4, help="untimed clock-ramp loop before the warm-up (0 under ncu)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded cpu_baseline sample")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_clip(workload, root, seed):
    from cama_b200 import synth
    if workload == "config3":
        spec = synth.config3_spec(seed=1 + seed, name=f"config3_s{seed}")
        spec.write_cama = False
        dataset = "nuscenes"
    elif workload == "config2_cama":
        spec = synth.config2_spec(seed=seed, name=f"config2cama_s{seed}")
        spec.write_nuscenes = False
        dataset = "cama"
    else:
        spec = synth.config2_spec(seed=seed, name=f"config2_s{seed}")
        spec.write_cama = False
        dataset = "nuscenes"
    return synth.write_clip(spec, root), dataset


def workload_name(workload):
    return {"config2": "BASELINE.json configs[1]: one clip, 40 frames x 6 cams, 200 polylines, 0.1 m densify (nuScenes-style labels)",
            "config2_cama": "configs[1] with CAMA labels: 40 frames x 6 cams, 200 polylines, 0.1 px densify (~1.0 M vertices)",
            "config3": "BASELINE.json configs[2]: site, 320 frames x 6 cams, 1600 polylines, 0.1 m densify"}[workload]


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while a region runs."""

    def __init__(self, index, period_s=0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._loaded = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                if self._loaded.is_set():
                    self.samples.append(mhz)
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for name, bit in names.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def load(self, on):
        (self._loaded.set if on else self._loaded.clear)()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(clip, dataset, budget_s):
    """Bounded sample of the workload on the host cores -> the cpu_baseline object."""
    from cama_b200 import synth
    from oracle.cpu_bench import CpuRunner, usable_cores
    runner = CpuRunner(synth.CAMA_CONFIGS, clip, dataset)
    try:
        n_frames = runner.n_frames
        probe_frames = max(1, min(2, n_frames))
        dt1, cf1 = runner.step("numpy", parallel=False, frames=probe_frames)      # as shipped: one process
        per_frame = dt1 / probe_frames
        # whole clip over all cores if it fits the budget, else as many frames as do
        est = per_frame * n_frames / max(runner.workers, 1) * 1.5
        frames = n_frames if est <= budget_s else max(runner.workers, int(budget_s / est * n_frames))
        runner.step("numpy", parallel=True, frames=min(frames, runner.workers))   # warm the pool
        dt, cf = runner.step("numpy", parallel=True, frames=frames)
        dtc, cfc = runner.step("c", parallel=True, frames=n_frames)
        dtc1, cfc1 = runner.step("c", parallel=False, frames=min(n_frames, 8))
        # the same loop with render_vectors (cama/dataset.py:119-126): camera JPEGs decoded, undistort-resized, drawn on
        img_frames = min(n_frames, runner.workers)
        synth.write_background_jpegs(clip, img_frames, seed=0)
        dti, cfi = runner.step("numpy_images", parallel=True, frames=img_frames)
        return {"value": cf / dt, "unit": UNIT, "cores": runner.workers, "kind": "port",
                "sample": f"{frames} of {n_frames} frames x {runner.n_cams} cams of the same clip, NumPy/OpenCV oracle "
                          f"(reference loop structure), frames sharded over {runner.workers} processes; blank backgrounds",
                "single_process": {"value": cf1 / dt1, "unit": UNIT, "cores": 1, "sample": f"{probe_frames} frames"},
                "with_images": {"value": cfi / dti, "unit": UNIT, "cores": runner.workers,
                                "sample": f"{img_frames} frames x {runner.n_cams} cams through render_vectors: cv2.imread of a 1600x900 JPEG, "
                                          "initUndistortRectifyMap + remap, then the same drawing (BASELINE.md's second CPU figure)"},
                "c_port": {"value": cfc / dtc, "unit": UNIT, "cores": runner.workers, "single_core": cfc1 / dtc1,
                           "note": "scalar C restatement (oracle/oracle.c), not what the reference runs"},
                "host": {"cpu": cpu_model(), "usable_cores": usable_cores(), "os_cpu_count": os.cpu_count()}}
    finally:
        runner.close()


def run_reference(args):
    """--impl reference: the reference's CPU path (NumPy/OpenCV port; the reference is pure Python
    and /root/reference does not exist on the GPU box) on all host cores."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from cama_b200 import synth
    from oracle.cpu_bench import CpuRunner
    with tempfile.TemporaryDirectory() as root:
        clip, dataset = make_clip(args.workload, root, 0)
        runner = CpuRunner(synth.CAMA_CONFIGS, clip, dataset)
        try:
            # bounded sample per step: the whole clip when it is small, else ~4 s worth of frames
            dt_probe, _ = runner.step("numpy", parallel=True, frames=runner.workers)
            per_step_frames = runner.n_frames
            est = dt_probe * runner.n_frames / runner.workers
            budget = 240.0 / max(args.steps + args.warmup, 1)
            if est > budget:
                per_step_frames = max(runner.workers, int(runner.n_frames * budget / est))
            for _ in range(args.warmup):
                runner.step("numpy", parallel=True, frames=per_step_frames)
            total_t, total_cf = 0.0, 0
            for _ in range(args.steps):
                dt, cf = runner.step("numpy", parallel=True, frames=per_step_frames)
                total_t += dt
                total_cf += cf
            value = total_cf / total_t
            sample = (f"{per_step_frames} of {runner.n_frames} frames x {runner.n_cams} cams per step, NumPy/OpenCV oracle "
                      f"(reference loop structure), frames sharded over {runner.workers} processes; blank backgrounds")
            line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
                    "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": workload_name(args.workload), "frames": runner.n_frames, "cams": runner.n_cams,
                               "cam_frames_per_step": per_step_frames * runner.n_cams},
                    "cpu_baseline": {"value": value, "unit": UNIT, "cores": runner.workers, "kind": "port", "sample": sample,
                                     "host": {"cpu": cpu_model(), "os_cpu_count": os.cpu_count()}},
                    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line), flush=True)
        finally:
            runner.close()


def config3_block(args, root, device):
    """The site of BASELINE.json configs[2] on one GPU: ms per step (one stream; the library pipelines its frame groups over
    its own lanes), per-phase times from CUDA events, raster and whole-step roofline figures."""
    import torch
    from cama_b200 import synth
    from cama_b200 import _native as N
    from cama_b200.batched import Reproject
    clip, dataset = make_clip("config3", root, 0)
    rp = Reproject(synth.CAMA_CONFIGS, clip, device=device)
    rt, res = rp.rt, rp.resident(dataset)
    idx, w2c = rp.frame_poses(dataset)
    F, C = len(idx), rp.renderer.n_cams
    w2c_dev = torch.from_numpy(w2c).to(rt.device)
    frames = torch.empty((F, C, H, W, 3), dtype=torch.uint8, device=rt.device)
    rp.renderer.render(res, w2c_dev, out=frames, mode=args.mode, check=True)
    stats = dict(rp.renderer.last_stats)
    steps = max(5, args.steps // 3)
    step = lambda: rp.renderer.render(res, w2c_dev, out=frames, mode=args.mode, check=False)
    for _ in range(3):
        step()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rt.profile_enable(steps)                       # (phase events: the frame groups then run in one pass on one stream)
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    phases = rt.profile_read().mean(axis=0)
    rt.profile_enable(0)
    frame_bytes, vertex_bytes = F * C * H * W * 3, F * 12 * res.n_vertices
    raster_ms = float(phases[3])
    return {"workload": workload_name("config3"), "frames": F, "cams": C, "vertices": res.n_vertices, "instances": res.n_instances,
            "value": F * C / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "records_total": int(stats["records_total"]), "record_capacity": int(stats["record_capacity"]),
            "roofline": {"bound": "hbm", "kernel": "binned_raster_kernel", "achieved": frame_bytes / (raster_ms * 1e-3) / 1e9, "unit": "GB/s",
                         "algorithmic_bytes_per_launch": int(frame_bytes), "launch_ms": raster_ms,
                         "launch_ms_source": "CUDA events on the launching stream around the raster phase, mean over a second run of the same steps",
                         "whole_step": {"algorithmic_bytes": int(frame_bytes + vertex_bytes), "achieved": (frame_bytes + vertex_bytes) / (ms * 1e-3) / 1e9},
                         "phase_ms": {name: float(phases[i]) for i, name in enumerate(N.PHASE_NAMES)}}}


def host_memory_probe(threads):
    """STREAM-like fill / copy bandwidth of this process's host threads (cama_host_bandwidth_probe), GB/s."""
    import ctypes
    from cama_b200 import _native as N
    fill, copy = ctypes.c_double(), ctypes.c_double()
    N.check(N.lib().cama_host_bandwidth_probe(512 << 20, int(threads), ctypes.byref(fill), ctypes.byref(copy)))
    return fill.value, copy.value


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    rank, local_rank, world = dist_env()
    assert world == 1, "several ranks run run_b200_sharded"
    tmp = tempfile.TemporaryDirectory()
    clip, dataset = make_clip(args.workload, tmp.name, rank)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(clip, dataset, args.cpu_seconds)        # before CUDA is initialised (forks workers)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    torch.cuda.set_device(local_rank)

    from cama_b200 import synth
    from cama_b200 import _native as N
    from cama_b200.batched import Reproject

    rp = Reproject(synth.CAMA_CONFIGS, clip, device=local_rank)
    rt = rp.rt
    res = rp.resident(dataset)
    idx, w2c_host = rp.frame_poses(dataset)
    F, C = len(idx), rp.renderer.n_cams
    cam_frames = F * C
    w2c_dev = torch.from_numpy(w2c_host).to(rt.device)
    frames = torch.empty((F, C, H, W, 3), dtype=torch.uint8, device=rt.device)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()

    def step(occupancy=(0, 0)):
        rp.renderer.geometry_ctas_per_sm, rp.renderer.raster_ctas_per_sm = occupancy
        rp.renderer.render(res, w2c_dev, out=frames, mode=args.mode, check=False)

    # Throughput mode: the K steps are dealt over `lanes` streams, each with its own workspace and output frames,
    # so that independent clips overlap (geometry of one under the store-bound raster of the other).
    n_lanes = max(1, args.lanes)
    lane_streams = [torch.cuda.Stream(device=rt.device) for _ in range(n_lanes)]
    lane_frames = [frames] + [torch.empty_like(frames) for _ in range(n_lanes - 1)]

    overlap_occupancy = (args.geometry_ctas, args.raster_ctas) if n_lanes > 1 else (0, 0)

    def lane_step(k):
        rp.renderer.geometry_ctas_per_sm, rp.renderer.raster_ctas_per_sm = overlap_occupancy
        lane = k % n_lanes
        with torch.cuda.stream(lane_streams[lane]):
            rp.renderer.render(res, w2c_dev, out=lane_frames[lane], mode=args.mode, check=False, lane=lane)

    def timed_lanes(steps):
        """-> ms for `steps` steps over the lanes, from an event on the main stream before the first launch to one
        after the last kernel of every lane."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for ls in lane_streams:
            ls.wait_event(e0)
        for k in range(steps):
            lane_step(k)
        for ls in lane_streams:
            done = torch.cuda.Event()
            done.record(ls)
            stream.wait_event(done)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    # sizing pass: settles the record-pool capacity (checked, synchronous) and verifies no overflow
    rp.renderer.render(res, w2c_dev, out=frames, mode=args.mode, check=True)
    stats = dict(rp.renderer.last_stats)
    assert not stats["overflow"]

    sampler = ClockSampler(local_rank)
    sampler.start()

    # untimed clock ramp (a 100-us step does not lift an idle GPU to its boost clock), then W warm-up steps
    t_end = time.perf_counter() + args.ramp_seconds
    while time.perf_counter() < t_end:
        step()
        torch.cuda.synchronize()
    for k in range(max(args.warmup, 3) * n_lanes):
        lane_step(k)
    torch.cuda.synchronize()
    for lf in lane_frames[1:]:
        assert torch.equal(lf, frames), "lanes disagree"

    # ---- device-resident throughput: exactly K steps, nothing else on the streams
    # (the kernels of a step are chained by programmatic dependent launch; events between them would serialise them)
    launches0 = rt.launches()
    sampler.load(True)
    ms_total = timed_lanes(args.steps)
    launches = rt.launches() - launches0
    # the same K steps on one stream (what a single clip sees)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step((0, 0))
    ev1.record(stream)
    barrier()
    ms_single = ev0.elapsed_time(ev1)

    # ---- the same K steps again with CUDA events recorded on the launching stream around every phase
    # (cama_ctx_profile_*): the per-kernel durations the roofline is computed from
    rt.profile_enable(args.steps)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record(stream)
    for _ in range(args.steps):
        step(overlap_occupancy)
    ev3.record(stream)
    barrier()
    sampler.load(False)
    ms_instrumented = ev2.elapsed_time(ev3)
    phases = rt.profile_read()                      # [K, 4] ms
    rt.profile_enable(0)
    rp.renderer.geometry_ctas_per_sm = rp.renderer.raster_ctas_per_sm = 0       # (library defaults for everything below)

    # ---- end to end through the public call: host poses -> finished frames in host memory.
    # Default transfer: the lit 8-pixel chunks cross PCIe and libcama_b200's host routine draws them into
    # the host frames (previous overlay blanked first); "dense" copies all frame bytes back instead.
    def e2e_loop(transfer, steps):
        # untimed: allocates host buffers (first touch of 373 MB of host frames), settles capacities, starts the helper
        # thread and the OpenMP team: at least W calls and 0.2 s
        t_warm = time.perf_counter()
        n_warm = 0
        while n_warm < max(args.warmup, 3) or time.perf_counter() - t_warm < 0.2:
            rp(dataset, mode=args.mode, transfer=transfer)
            n_warm += 1
        barrier()
        sampler.load(True)
        t0 = time.perf_counter()
        for _ in range(steps):
            _, host_frames = rp(dataset, mode=args.mode, transfer=transfer)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sampler.load(False)
        return dt, int(host_frames[:, :, ::9, ::9].sum()), dict(rp.last_transfer)

    e2e_steps = max(3, min(args.steps, 20))
    e2e_s, checksum, transfer = e2e_loop("sparse", e2e_steps)
    dense_s, checksum_dense, transfer_dense = e2e_loop("dense", max(3, e2e_steps // 2))
    dense_steps = max(3, e2e_steps // 2)
    assert checksum == checksum_dense, "sparse and dense transfers disagree"
    host_fill_gbs, host_copy_gbs = host_memory_probe(rp.host_threads)

    # ---- the zero-change drop-in path: the three calls unmodified main.py makes per frame (main.py:57-59), host lists of
    # NumPy arrays in and out as the reference's protocol demands, one process
    from tools.dropin_bench import dropin_loop
    from cama_b200.dataset import ClipManager
    cm_dropin = ClipManager(synth.CAMA_CONFIGS, clip, device=local_rank, progress=False)
    dropin_loop(cm_dropin, dataset, H, W, max_frames=3)
    torch.cuda.synchronize()
    dropin_t, dropin_done, _ = dropin_loop(cm_dropin, dataset, H, W, max_frames=min(F, 20))
    dropin_s = sum(dropin_t.values())

    # ---- BASELINE.json configs[2] on this GPU: the site (320 frames x 6 cameras, ~767 k vertices), whose 5.9 GB of algorithmic
    # traffic per step is the size BASELINE.md asks the roofline fraction for (config 2 moves 0.42 GB: a latency-sized problem)
    site = None
    if not args.no_config3 and args.workload == "config2":
        site = config3_block(args, tmp.name, local_rank)

    clocks = sampler.stop()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None                   # dram bytes of the raster kernel from the committed ncu --set full capture
        try:
            with open(os.path.join(REPO, "profiles", "raster_traffic.json")) as fh:
                traffic = float(json.load(fh)["traffic_bytes"]) * cam_frames / 240.0
        except (OSError, KeyError, ValueError):
            pass
        raster_ms = float(np.mean(phases[:, 3]))
        frame_bytes = cam_frames * H * W * 3
        vertex_bytes = F * 12 * res.n_vertices
        achieved = frame_bytes / (raster_ms * 1e-3) / 1e9
        ms_per_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": world * cam_frames / (ms_per_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload),
                       "frames": F, "cams": C, "cam_frames_per_step_per_gpu": cam_frames, "vertices": res.n_vertices,
                       "instances": res.n_instances, "raster_mode": {1: "plane", 2: "binned"}[stats["mode"]],
                       "l2": f"no flush needed: every step writes {frame_bytes / 1e6:.0f} MB of frames (> 126 MB L2); "
                             f"the {res.n_vertices * 16 / 1e6:.1f} MB vertex array is L2-resident by nature (re-read for each of the {F} frames)",
                       "background": "blank (black) frames, as in the reference CPU timing",
                       "streams": n_lanes, "geometry_ctas_per_sm": overlap_occupancy[0] or 4, "raster_ctas_per_sm": overlap_occupancy[1] or 4,
                       "streams_note": f"the K steps are dealt over {n_lanes} CUDA streams with separate workspaces and output frames (independent clips "
                                       "overlap); single_stream = the same K steps back to back on one stream"},
            "single_stream": {"value": world * cam_frames / (ms_single / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms_single / args.steps},
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
            "e2e": {"value": world * cam_frames * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(w2c_host.nbytes), "d2h_bytes_per_step": int(transfer["d2h_bytes"]) + 48,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "call": "cama_b200.batched.Reproject.__call__(dataset): host pose seek + float32 inverse, H2D of the poses from pinned "
                            "memory, cama_clip_render (sparse output), D2H of the lit 8-pixel chunk records into pinned memory in slices, "
                            "cama_overlay_apply_host blanks the previous overlay (helper thread, during the former) and draws the new one "
                            "into the host frames [F,C,540,960,3]",
                    "transfer": "sparse", "overlay_records": int(transfer["records"]), "checksum": checksum,
                    "host_memory": {"fill_gbs": host_fill_gbs, "copy_gbs": host_copy_gbs, "threads": int(rp.host_threads),
                                    "line_traffic_bytes_per_step": [int(transfer["records"]) * 2 * 48, int(transfer["records"]) * 2 * 128],
                                    "note": "fill/copy: STREAM-like figures of the worker pool that blanks and draws the lit chunks; line_traffic = "
                                            "records x 2 (blank + draw) x (read-for-ownership + write-back) of the 64-byte lines touched, between 24/64 of a "
                                            "line per chunk (adjacent chunks share lines) and a line per chunk — what the call moves to and from DRAM if no "
                                            "lit line stays in the last-level cache between blank and draw (with one clip's ~80 MB of lit lines most do)"},
                    "host_draw_threads": int(rp.host_threads), "host_cores": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
                    "dense": {"value": world * cam_frames * dense_steps / dense_s, "unit": UNIT, "ms_per_step": 1e3 * dense_s / dense_steps,
                              "d2h_bytes_per_step": int(frame_bytes), "d2h_gbs": frame_bytes * dense_steps / dense_s / 1e9,
                              "note": "same call with transfer='dense': all frame bytes rendered in HBM and copied back (PCIe-bound)"}},
            "dropin": {"value": dropin_done / dropin_s, "unit": UNIT, "cam_frames": dropin_done,
                       "ms_per_frame": {k: 1e3 * v / (dropin_done / C) for k, v in dropin_t.items()},
                       "call": "unmodified main.py's per-frame protocol on one process: ClipManager.yield_frame -> project_all_camera -> "
                               "CameraManager.render_maps on blank frames, host lists of NumPy arrays in and out (one device operator per call; "
                               "the clip's vertices stay resident, the cropped points are handed from yield_frame to project_all_camera on the device)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "binned_raster_kernel" if stats["mode"] == 2 else "plane_raster_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "traffic": traffic, "traffic_source": "profiles/raster_traffic.json (ncu --set full, config 2, scaled by cam-frames)",
                         "algorithmic_bytes_per_launch": int(frame_bytes), "launch_ms": raster_ms,
                         "launch_ms_source": f"CUDA events on the launching stream around the kernel, mean over a second run of the same {args.steps} "
                                             f"steps ({ms_instrumented / args.steps:.4f} ms per step with the phase events in the stream)",
                         "whole_step": {"algorithmic_bytes": int(frame_bytes + vertex_bytes),
                                        "achieved": (frame_bytes + vertex_bytes) / (ms_per_step * 1e-3) / 1e9,
                                        "frac": (frame_bytes + vertex_bytes) / (ms_per_step * 1e-3) / 1e9 / peak},
                         "phase_ms": {name: float(np.mean(phases[:, i])) for i, name in enumerate(N.PHASE_NAMES)}},
            "records": {k: int(stats[k]) for k in ("records_total", "record_capacity_needed", "record_capacity")},
        }
        if site is not None:
            site["roofline"]["peak"] = peak
            site["roofline"]["frac"] = site["roofline"]["achieved"] / peak
            site["roofline"]["whole_step"]["frac"] = site["roofline"]["whole_step"]["achieved"] / peak
            line["config3"] = site
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    tmp.cleanup()


# ---------------------------------------------------------------------------------------------- GPU arm, N > 1
def run_b200_sharded(args):
    """BASELINE.json configs[3]: one site, its frames sharded over the ranks, every frame assembled on every rank."""
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    tmp = tempfile.TemporaryDirectory()
    clip, dataset = make_clip(args.workload, tmp.name, 0)              # the same site on every rank (seeded)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cama_b200 import shard, synth
    from cama_b200 import _native as N
    from cama_b200.batched import Reproject

    rp = Reproject(synth.CAMA_CONFIGS, clip, device=local_rank)
    rt, r = rp.rt, rp.renderer
    res = rp.resident(dataset)
    idx, w2c_host = rp.frame_poses(dataset)
    F, C = len(idx), r.n_cams
    cam_frames = F * C
    lo, hi = shard.frame_block(F, rank, world)
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=rt.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm):
        """-> ms per step: events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    asm = shard.SiteAssembler(rp, dataset, mode=args.mode)             # collective: sizes the slots, maps the peers' mailboxes
    frames_all = torch.empty((F, C, H, W, 3), dtype=torch.uint8, device=rt.device)
    local = torch.empty((hi - lo, C, H, W, 3), dtype=torch.uint8, device=rt.device)
    w2c_block = asm.w2c_dev
    r.render(res, w2c_block, out=local, mode=args.mode, check=True)    # sizing pass of the dense block render
    center_records = (r.last_stats or {}).get("records_total", 0)      # centre records of this rank's block

    def step_assembled():
        if asm.available:
            asm.step(out=frames_all)
        else:                                                          # no peer-to-peer path: NCCL all-gather of the records
            records, n, fmt = r.render_overlay(res, w2c_block, mode=args.mode)
            everyone, counts = shard.gather_records(records, n)
            for peer in range(world):
                p_lo, p_hi = shard.frame_block(F, peer, world)
                if p_hi > p_lo:
                    r.expand_overlay(everyone[peer], counts[peer], fmt, res.palette, p_hi - p_lo, out=frames_all[p_lo:p_hi])

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_end = time.perf_counter() + args.ramp_seconds
    while time.perf_counter() < t_end:
        r.render(res, w2c_block, out=local, mode=args.mode, check=False)
        torch.cuda.synchronize()
    warm = max(args.warmup, 3)

    # ---- the headline: K steps, each = render of the rank's block + exchange + all frames rebuilt on every rank
    launches0 = rt.launches()
    sampler.load(True)
    ms_step = timed(step_assembled, args.steps, warm)
    launches = (rt.launches() - launches0) // max(1, args.steps + warm) * args.steps
    status = asm.exchange.status_code() if asm.available else 0
    # the assembly alone — zero-fill + expand of the records of the last step: every frame byte of the site written once
    # on this rank (the dominant kernels of the step) — on the launching stream; then one more full step (the frames
    # are compared below)
    if asm.available and asm.kind == "lists":
        ms_assembly = timed(lambda: asm.exchange.reraster(r, res, frames_all), max(5, args.steps // 3), 2)
    elif asm.available:
        ms_assembly = timed(lambda: asm.exchange.reassemble(r, res, F, frames_all), max(5, args.steps // 3), 2)
    else:
        ms_assembly = timed(lambda: N.check(N.lib().cama_frames_clear(rt.ctx, rt.ptr(frames_all), frames_all.numel(), rt.stream())), max(5, args.steps // 3), 2)
    step_assembled()
    torch.cuda.synchronize()
    # ---- compute only: every rank renders its block densely, nothing is exchanged
    ms_compute = timed(lambda: r.render(res, w2c_block, out=local, mode=args.mode, check=False), args.steps, warm)
    # ---- north_star's literal collective: dense render + one NCCL all-gather of the uint8 frames
    dense_steps = max(3, args.steps // 10)
    gathered = torch.empty((world * shard.block_size(F, world), C, H, W, 3), dtype=torch.uint8, device=rt.device)
    padded = local if hi - lo == shard.block_size(F, world) else torch.zeros((shard.block_size(F, world), C, H, W, 3), dtype=torch.uint8, device=rt.device)

    def step_dense():
        r.render(res, w2c_block, out=padded[:hi - lo], mode=args.mode, check=False)
        dist.all_gather_into_tensor(gathered, padded)

    ms_dense = timed(step_dense, dense_steps, 2)
    sampler.load(False)
    dense_equal = bool(torch.equal(gathered[:F], frames_all))
    del gathered
    # ---- one GPU, same workload: every rank renders the whole site alone (rank 0's time is reported)
    w2c_all = torch.from_numpy(w2c_host).to(rt.device)
    whole = torch.empty((F, C, H, W, 3), dtype=torch.uint8, device=rt.device)
    r.render(res, w2c_all, out=whole, mode=args.mode, check=True)
    single_steps = max(3, args.steps // 5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(single_steps):
        r.render(res, w2c_all, out=whole, mode=args.mode, check=False)
    e1.record(stream)
    barrier()
    ms_single_gpu = e0.elapsed_time(e1) / single_steps
    # ---- verification inside the bench: the assembled frames of EVERY rank equal the one-GPU render of the site
    equal = torch.tensor([int(torch.equal(frames_all, whole)), int(dense_equal)], dtype=torch.int32, device=rt.device)
    dist.all_reduce(equal, op=dist.ReduceOp.MIN)
    checksum = int(frames_all[:, :, ::9, ::9].sum().item())
    del whole

    # ---- end to end through the public call, frames in HOST memory: every rank takes its frame block through
    # Reproject.__call__ (host pose seek + float32 inverse, H2D of the poses, sparse render, records D2H, host draw)
    def e2e_loop(steps):
        n_warm, t_warm = 0, time.perf_counter()
        while n_warm < warm or time.perf_counter() - t_warm < 0.2:
            rp(dataset, mode=args.mode, frame_range=(lo, hi))
            n_warm += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            _, host_frames = rp(dataset, mode=args.mode, frame_range=(lo, hi))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, int(host_frames[:, :, ::9, ::9].sum()), dict(rp.last_transfer)

    e2e_steps = max(3, min(args.steps, 20))
    e2e_s, e2e_sum, transfer = e2e_loop(e2e_steps)
    e2e_s = max_over_ranks(e2e_s)
    sums = torch.tensor([e2e_sum, int(transfer["records"])], dtype=torch.int64, device=rt.device)
    dist.all_reduce(sums)
    barrier()                                                          # every rank probes at the same time: the contended figure
    host_bw = torch.tensor(host_memory_probe(rp.host_threads), dtype=torch.float64, device=rt.device)
    dist.all_reduce(host_bw)
    clocks = sampler.stop()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        frame_bytes = cam_frames * H * W * 3                          # written once on EVERY rank
        vertex_bytes = (hi - lo) * 12 * res.n_vertices
        records = int(transfer["records"])
        line = {
            "metric": METRIC, "value": cam_frames / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload) + f"; frames sharded over {world} GPUs in contiguous blocks (BASELINE.json configs[3]), "
                                   "all frames assembled on every GPU inside the timed region",
                       "frames": F, "cams": C, "cam_frames_per_step": cam_frames, "frames_per_gpu": hi - lo, "vertices": res.n_vertices,
                       "instances": res.n_instances,
                       "assembly": ("peer memory, centre records: every rank's geometry kernel appends the records of its frame block to the per-band lists "
                                    "of EVERY rank (peer stores over NVLink while it runs), cama_peer_publish_cursors / cama_peer_wait hand the lists "
                                    "over, every rank rasters all frames (no zero-fill, no expand, no NCCL call, no host round trip in the step)")
                                   if asm.available and asm.kind == "lists" else
                                   ("peer memory, lit chunks: the raster mirrors its lit-chunk records into every peer's mailbox over NVLink while it runs, "
                                    "cama_peer_publish / cama_peer_expand into zero-filled frames (no NCCL call and no host round trip in the step)")
                                   if asm.available else f"NCCL all-gather of the lit-chunk records + cama_overlay_expand (no peer-to-peer path: {asm.exchange.error})",
                       "l2": f"no flush needed: every step writes {frame_bytes / 1e9:.2f} GB of frames per GPU (> 126 MB L2)",
                       "background": "blank (black) frames, as in the reference CPU timing"},
            "verified": {"assembled_equals_single_gpu_render_on_every_rank": bool(equal[0].item()),
                         "dense_allgather_equals_assembled_on_every_rank": bool(equal[1].item()),
                         "peer_status": status, "checksum": checksum},
            "single_gpu_same_workload": {"value": cam_frames / (ms_single_gpu * 1e-3), "unit": UNIT, "ms_per_step": ms_single_gpu,
                                         "note": "rank 0 renders all frames of the site alone, frames left in its HBM (what the assembled result is compared with)"},
            "compute_only": {"value": cam_frames / (ms_compute * 1e-3), "unit": UNIT, "ms_per_step": ms_compute,
                             "note": "every rank renders its frame block densely, nothing exchanged"},
            "dense_allgather": {"value": cam_frames / (ms_dense * 1e-3), "unit": UNIT, "ms_per_step": ms_dense,
                                "bytes_received_per_gpu": int(frame_bytes * (world - 1) / world),
                                "nvlink_gbs_in": frame_bytes * (world - 1) / world / (ms_dense * 1e-3) / 1e9,
                                "note": "dense render + ONE NCCL all_gather_into_tensor of the uint8 frames (north_star's literal collective): NVLink-bound"},
            "exchange": ({"kind": "centre records (4 B)", "records_per_gpu": int(center_records), "bytes_sent_per_gpu": int(center_records) * 4 * (world - 1),
                          "bytes_received_per_gpu": None} if asm.available and asm.kind == "lists" else
                         {"kind": "lit-chunk records (12 B)", "records_per_gpu": records, "bytes_sent_per_gpu": records * 12 * (world - 1), "bytes_received_per_gpu": None}),
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
            "e2e": {"value": cam_frames * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(w2c_host[lo:hi].nbytes) * world,
                    "d2h_bytes_per_step": (int(transfer["d2h_bytes"]) + 48) * world, "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "call": "every rank: cama_b200.batched.Reproject.__call__(dataset, frame_range=its block): host pose seek + float32 inverse, "
                            "H2D of the poses from pinned memory, cama_clip_render (sparse output), D2H of the lit-chunk records, "
                            "cama_overlay_apply_host into the rank's host frames; the site's frames end up in host memory once, split over the ranks",
                    "checksum_all_ranks": int(sums[0].item()), "host_draw_threads_per_rank": int(rp.host_threads),
                    "host_memory": {"fill_gbs_all_ranks": float(host_bw[0].item()), "copy_gbs_all_ranks": float(host_bw[1].item()),
                                    "line_traffic_bytes_per_step": [int(sums[1].item()) * 2 * 48, int(sums[1].item()) * 2 * 128],
                                    "ms_per_step_at_copy_bandwidth": [int(sums[1].item()) * 2 * b / max(float(host_bw[1].item()), 1e-9) / 1e6 for b in (48, 128)],
                                    "note": "fill/copy: STREAM-like figures of all ranks' worker pools running at once (they share the box's memory system); "
                                            "line_traffic = records x 2 (blank + draw) x (read-for-ownership + write-back) of the 64-byte lines touched, between "
                                            "24/64 of a line per chunk (adjacent chunks share lines) and a line per chunk; with N ranks x 373 MB of host frames "
                                            "the lit lines do not stay in the last-level cache, so the step is bound by this traffic at the bandwidth these "
                                            "threads reach (ms_per_step_at_copy_bandwidth: the two ends)"},
                    "host_cores": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": ("binned_raster_kernel over all frames of the site (every frame byte written once per rank)" if asm.available and asm.kind == "lists"
                                                    else "frames_clear_kernel + peer_expand_kernel (the assembly: every frame byte of the site written once per rank, then the lit chunks)"),
                         "achieved": frame_bytes / (ms_assembly * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": frame_bytes / (ms_assembly * 1e-3) / 1e9 / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "traffic": None, "algorithmic_bytes_per_launch": int(frame_bytes), "launch_ms": ms_assembly,
                         "launch_ms_source": "CUDA events on the launching stream around the assembly of the last step's records alone (raster phase over all "
                                             "frames, or zero-fill + expand; no geometry, no exchange; includes the ~10 us of prep + classify), barrier on both "
                                             "sides, max over ranks",
                         "whole_step": {"algorithmic_bytes": int(frame_bytes + vertex_bytes),
                                        "achieved": (frame_bytes + vertex_bytes) / (ms_step * 1e-3) / 1e9,
                                        "frac": (frame_bytes + vertex_bytes) / (ms_step * 1e-3) / 1e9 / peak,
                                        "note": "per rank: all frames of the site written once + the rank's share of the vertex reads, over the whole step"}},
        }
        line["exchange"]["bytes_received_per_gpu"] = line["exchange"]["bytes_sent_per_gpu"]
        print(json.dumps(line), flush=True)
    dist.barrier()
    if asm.available:
        asm.exchange.close()
    dist.destroy_process_group()
    tmp.cleanup()


def main():
    args = parse_args()
    _, _, world = dist_env()
    if args.workload is None:
        args.workload = "config2" if max(world, args.gpus) == 1 else "config3"
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus} ...")
    elif world > 1:
        run_b200_sharded(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
