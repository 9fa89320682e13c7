"""CPU timing of the oracle — TEST/BENCH INFRASTRUCTURE, NOT PRODUCT.

Used only by bench.py (``cpu_baseline`` leg and ``--impl reference`` arm).  Times the NumPy /
OpenCV restatement of the reference's frame loop (oracle/cama_oracle.py — same structure and
the same per-point ``cv2.circle`` calls as /root/reference/cama/dataset.py:78-126 +
cama/reproject.py:108-131,187-205,246-257) on a clip directory:

* ``single``  one Python process, the reference exactly as shipped (it is single-threaded);
* ``pool``    the frames of the clip sharded over worker processes, one ClipOracle each — the
              box-level figure "with all the host threads it can use" (frames are independent).

The plain-C restatement (oracle/oracle.c) is timed the same way as a second, much stronger CPU
figure; it is *not* what the reference does, it is what a scalar C rewrite would do.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:                      # pragma: no cover
        return os.cpu_count() or 1


def _init(configs, clip_path, dataset):
    import cv2
    cv2.setNumThreads(1)                         # cv2.circle is not threaded anyway; keep workers from oversubscribing
    from oracle import cama_oracle as orc
    oc = orc.ClipOracle(configs, clip_path)
    poses = oc.world_to_chassis_per_frame(dataset)
    flat, classes, counts = orc.flatten(oc.instance_maps[dataset], 3)
    bgr = np.array([orc.CLASS_RGB["lane_marking" if c == "lane_marking" else "Crosswalk_Line"][::-1] for c in classes], np.uint8)
    _STATE.update(oc=oc, orc=orc, dataset=dataset, poses=poses, flat=flat,
                  offs=np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), bgr=bgr)


def _numpy_frames(span):
    """The reference loop (NumPy + cv2.circle) for frames [lo, hi) -> (#cam-frames, checksum)."""
    lo, hi = span
    oc, orc = _STATE["oc"], _STATE["orc"]
    h, w = orc.OUTPUT_HW
    done, check = 0, 0
    for _, w2c in _STATE["poses"][lo:hi]:
        chassis = orc.crop_instances(orc.transform_instances(oc.instance_maps[_STATE["dataset"]], w2c))
        per_cam = oc.project_all(chassis)
        for cam in oc.cameras:
            img = orc.render_instances(np.zeros((h, w, 3), np.uint8), per_cam[cam])
            check += int(img[::7, ::7].sum())
            done += 1
    return done, check


def _numpy_frames_images(span):
    """The same with render_vectors (cama/dataset.py:119-126): every camera image is read from its JPEG, undistort-
    resized and drawn on in place — the second CPU figure BASELINE.md asks for."""
    lo, hi = span
    oc, orc = _STATE["oc"], _STATE["orc"]
    done, check = 0, 0
    for image_idx, w2c in _STATE["poses"][lo:hi]:
        chassis = orc.crop_instances(orc.transform_instances(oc.instance_maps[_STATE["dataset"]], w2c))
        images = oc.render_vectors(oc.project_all(chassis), image_idx)
        for cam in oc.cameras:
            check += int(images[cam][::7, ::7].sum())
            done += 1
    return done, check


def _c_frames(span):
    lo, hi = span
    from oracle import oracle_c
    oc, orc = _STATE["oc"], _STATE["orc"]
    h, w = orc.OUTPUT_HW
    w2c = np.stack([m for _, m in _STATE["poses"][lo:hi]]) if hi > lo else np.zeros((0, 4, 4), np.float32)
    box = [orc.CROP_BOX[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]
    frames, _, _ = oracle_c.clip_render(_STATE["flat"], _STATE["offs"], _STATE["bgr"], w2c, np.stack(oc.chassis2cam),
                                        np.stack(oc.K), box, h, w, want_counts=False)
    return frames.shape[0] * frames.shape[1], int(frames[:, :, ::7, ::7].sum())


def _spans(n_frames, parts):
    edges = np.linspace(0, n_frames, parts + 1).astype(int)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


class CpuRunner:
    """Persistent worker pool; ``step(kind)`` renders every frame of the clip once."""

    def __init__(self, configs, clip_path, dataset, workers=None):
        self.workers = max(1, workers or usable_cores())
        _init(configs, clip_path, dataset)
        self.n_frames = len(_STATE["poses"])
        self.n_cams = len(_STATE["oc"].cameras)
        self.workers = min(self.workers, max(self.n_frames, 1))
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.workers, initializer=_init, initargs=(configs, clip_path, dataset)) if self.workers > 1 else None

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def step(self, kind="numpy", parallel=True, frames=None):
        """-> (seconds, cam-frames rendered).  ``frames`` limits the sample to the first n frames."""
        fn = {"numpy": _numpy_frames, "numpy_images": _numpy_frames_images, "c": _c_frames}[kind]
        n = self.n_frames if frames is None else min(frames, self.n_frames)
        t0 = time.perf_counter()
        if parallel and self.pool is not None:
            results = self.pool.map(fn, _spans(n, self.workers), chunksize=1)
        else:
            results = [fn((0, n))]
        dt = time.perf_counter() - t0
        return dt, sum(r[0] for r in results)
