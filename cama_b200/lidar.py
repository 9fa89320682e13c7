"""LiDAR aggregation (BASELINE.json configs[4]): every sweep of a clip moved
lidar -> chassis -> world with its pose and counted into one voxel grid, in one launch.

What the reference snapshot provides is reused as it is — the sweep reader ``DatasetReader.yield_lidar``
(cama/dataset_reader.py:45-51), the calibration graph ``get_extrinsic`` (:222-248), the chassis->world trajectory
of ``ClipManager.get_pt_nuscenes`` (cama/dataset.py:71-76) and ``PoseTransformer.seek_by_timestamp``
(cama/pose_transformer.py:589-652) — all on the host, exactly like the camera path.  The voxel accumulation is
NOT in the snapshot (camav2 branch, README.md:17-20): its definition is this package's own (include/cama_b200.h,
``cama_lidar_accumulate``) and ``oracle/lidar_oracle.py`` restates it — parity unpinned.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N
from .runtime import get_runtime

DEFAULT_GRID = {"origin": (-100.0, -100.0, -5.0), "voxel": (0.25, 0.25, 0.25), "dims": (800, 800, 40)}


def sweep_transforms(pose_transformer, lidar2chassis, stamps, t_max_diff=0.5):
    """-> (kept sweep indices, float64 [k,4,4] lidar->world): ``chassis2world(t) @ lidar2chassis`` per sweep, the
    pose sought like ``ClipManager.yield_frame`` does for camera frames (interpolated, a ``RuntimeError`` skips
    the sweep: cama/dataset.py:90-96)."""
    kept, mats = [], []
    lidar2chassis = np.asarray(lidar2chassis, dtype=np.float64)
    for i, t in enumerate(stamps):
        try:
            chassis2world = pose_transformer.seek_by_timestamp(float(t), t_max_diff=t_max_diff, interpolate=True)
        except RuntimeError:
            continue
        kept.append(i)
        mats.append(np.asarray(chassis2world, dtype=np.float64) @ lidar2chassis)
    return kept, (np.stack(mats) if mats else np.zeros((0, 4, 4)))


class LidarAggregator:
    """Voxel-count aggregation of LiDAR sweeps on one GPU."""

    def __init__(self, grid=None, device=None):
        g = dict(DEFAULT_GRID if grid is None else grid)
        self.origin = tuple(float(v) for v in g["origin"])
        self.voxel = tuple(float(v) for v in g["voxel"])
        self.dims = tuple(int(v) for v in g["dims"])                      # nx, ny, nz
        self.rt = get_runtime(device)

    def _grid(self):
        g = N.VoxelGrid()
        g.origin = (ctypes.c_double * 3)(*self.origin)
        g.voxel = (ctypes.c_double * 3)(*self.voxel)
        g.dims = (ctypes.c_int32 * 3)(*self.dims)
        return g

    def new_counts(self):
        import torch
        nx, ny, nz = self.dims
        return torch.zeros((nz, ny, nx), dtype=torch.int32, device=self.rt.device)      # uint32 bit pattern

    def accumulate(self, sweeps, transforms, counts=None):
        """sweeps: list of (n_i, >=3) float64 arrays (x y z first); transforms: float64 [S,4,4] lidar->world.
        -> (counts torch int32 [nz,ny,nx] on the device (uint32 bit pattern), number of points counted)."""
        import torch
        rt = self.rt
        assert len(sweeps) == len(transforms)
        if counts is None:
            counts = self.new_counts()
        if len(sweeps) == 0:
            return counts, 0
        arrays = [np.asarray(s, dtype=np.float64) for s in sweeps]
        widths = {a.reshape(len(a), -1).shape[1] for a in arrays if a.size}
        assert len(widths) <= 1 and all(w >= 3 for w in widths), "sweeps must share one row width >= 3"
        width = widths.pop() if widths else 6
        rows = [np.ascontiguousarray(a.reshape(len(a), width)) if a.size else np.zeros((0, width)) for a in arrays]
        offsets = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        flat = np.concatenate(rows, axis=0) if offsets[-1] else np.zeros((0, width))
        return self.accumulate_device(rt.to_device(flat), rt.to_device(offsets), rt.to_device(np.ascontiguousarray(transforms, dtype=np.float64).reshape(-1, 16)),
                                      counts)

    def accumulate_device(self, points_dev, offsets_dev, transforms_dev, counts=None, read_inside=True):
        """Device-resident inputs: points float64 [N,w], offsets int64 [S+1], transforms float64 [S,16]."""
        import torch
        rt = self.rt
        if counts is None:
            counts = self.new_counts()
        n_sweeps = int(offsets_dev.shape[0]) - 1
        inside = torch.zeros(1, dtype=torch.int64, device=rt.device) if read_inside else None
        grid = self._grid()
        N.check(N.lib().cama_lidar_accumulate(rt.ctx, rt.ptr(points_dev), int(points_dev.shape[1]) if points_dev.ndim == 2 else 6, rt.ptr(offsets_dev), n_sweeps,
                                              rt.ptr(transforms_dev), ctypes.byref(grid), rt.ptr(counts), rt.ptr(inside) if inside is not None else None,
                                              rt.stream()))
        return counts, (int(inside.item()) if inside is not None else None)

    def aggregate_clip(self, clip_manager, dataset="nuscenes", start_idx=None, end_idx=None, rank=0, world_size=1):
        """All sweeps of ``clip_manager``'s clip (or this rank's contiguous block of them): reads the .bin files,
        looks the poses up, accumulates.  -> (counts, points counted, sweeps used)."""
        from .dataset_reader import DatasetReader
        from .shard import frame_block
        dr = DatasetReader(clip_manager.clip_path)
        lidar2chassis = dr.get_extrinsic("lidar_top", "chassis")
        pt, _ = clip_manager._trajectory(dataset)
        stamped = list(dr.yield_lidar(start_idx=start_idx, end_idx=end_idx))
        lo, hi = frame_block(len(stamped), rank, world_size)
        stamped = stamped[lo:hi]
        kept, mats = sweep_transforms(pt, lidar2chassis, [t for t, _ in stamped])
        counts, inside = self.accumulate([stamped[i][1] for i in kept], mats)
        return counts, inside, len(kept)


def allreduce_counts(counts, group=None):
    """Sum of the ranks' voxel counts on every rank (sweeps are sharded across ranks; NCCL on GPUs, gloo in the CPU
    tests).  int32 addition of the uint32 bit patterns is the uint32 sum modulo 2^32."""
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
